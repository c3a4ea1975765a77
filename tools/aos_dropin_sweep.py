#!/usr/bin/env python
"""gpu_transport_photons drop-in (bgpu_transport_photons_aos) on the bench workload for several numbers of host copy
threads: python tools/aos_dropin_sweep.py [--copiers 1,2,4,6,8] [--photons 10000000]"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from branson_b200 import driver  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--copiers", default="1,2,4,6,8")
ap.add_argument("--photons", type=int, default=bench.PHOTONS_PER_GPU)
ap.add_argument("--slices", default="", help="comma list of minimum slice sizes in photons (BRANSON_AOS_SLICE)")
args = ap.parse_args()
deck = bench.make_deck(1, 8, args.photons)  # the run must not be over when the drop-in is timed
xml = deck.write(os.path.join(tempfile.mkdtemp(prefix="aos_"), "deck.xml"))
d = driver.Driver(xml, n_groups=bench.N_GROUPS, device=0, mesh_on_device=False)
for _ in range(3):
    d.cycle()
for sl in (args.slices.split(",") if args.slices else [""]):
    if sl:
        os.environ["BRANSON_AOS_SLICE"] = sl
    for n in args.copiers.split(","):
        os.environ["BRANSON_AOS_COPIERS"] = n
        r = bench.aos_dropin_block(d, bench.N_GROUPS, 0, repeats=3)
        print(f"slice >= {sl or 'default'} copiers {n}: {r['ms']:8.2f} ms  {r['value'] / 1e6:8.2f} M histories/s  "
              f"({r['photons']} photons)", flush=True)
d.close()
