#!/usr/bin/env python
"""Host-side phase times of the cycle driver on one GPU (device mesh): python tools/cycle_phases.py [deck] [photons]
deck: big_cube (default) | hohlraum_single | hohlraum_multi | hot_zone"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from branson_b200 import decks, driver  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "big_cube"
mk = {"big_cube": lambda p: decks.big_cube(n=200, photons=p or 125_000_000, t_stop=0.003),
      "hohlraum_single": lambda p: decks.hohlraum_single(photons=p or 10_000_000, t_stop=0.04),
      "hohlraum_multi": lambda p: decks.hohlraum_multi(photons=p or 31_250_000, t_stop=0.004),
      "hot_zone": lambda p: decks.hot_zone(photons=p or 10_000_000, t_stop=0.04)}[name]
deck = mk(int(float(sys.argv[2])) if len(sys.argv) > 2 else 0)
xml = deck.write(os.path.join(tempfile.mkdtemp(), "deck.xml"))
d = driver.Driver(xml, n_groups=deck.n_groups, device=0, mesh_on_device=True)
while not d.finished():
    t0 = time.perf_counter()
    r = d.cycle()
    t1 = time.perf_counter()
    d.array("T_e")
    t2 = time.perf_counter()
    print({k: round(1e3 * v, 2) for k, v in r.items() if k.startswith("t_")},
          "cycle() %.1f ms, T_e read-back %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)),
          {k: round(r["gpu"][k], 2) for k in ("ms_source", "ms_transport", "ms_census")}, flush=True)
d.close()
