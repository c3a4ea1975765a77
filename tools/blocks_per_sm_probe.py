"""Resident CTAs per SM on the two small decks (1e6 photons: 1.4 ms launches whose tail is a visible share): the occupancy
maximum (0 = auto: 5) is the best setting -- marshak 753 / 588 / 707 / 748 M histories/s at auto / 2 / 3 / 4 CTAs per SM,
hot_zone 716 / 548 / 655 / 703 (round 2, one B200).  python tools/blocks_per_sm_probe.py"""
import sys, os, tempfile
sys.path.insert(0, ".")
from branson_b200 import decks, driver
for name, deck, cyc in (("marshak", decks.marshak_wave(t_stop=0.08), 8), ("hot_zone", decks.hot_zone(t_stop=0.08), 8)):
    for bps in (0, 2, 3, 4):
        d = driver.Driver(deck.write(os.path.join(tempfile.mkdtemp(), "d.xml")), n_groups=deck.n_groups, device=0, mesh_on_device=True)
        d.gpu_context().set_launch(blocks_per_sm=bps)
        ms = []
        for c in range(cyc):
            g = d.cycle()["gpu"]; ms.append((g["n_transported"], g["ms_transport"], g["transport_kernel"]))
        d.close()
        n, t, k = ms[-1]
        print(f"{name} blocks_per_sm={bps}: last cycle {n/t/1e3:.1f} M/s ({t:.3f} ms, kernel {k}); mean of last 4: {sum(x[0] for x in ms[-4:])/sum(x[1] for x in ms[-4:])/1e3:.1f}")
