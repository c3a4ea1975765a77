import math, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from branson_b200 import decks, driver, gpu
deck = decks.hohlraum_single(t_stop=0.02)
for mode in (gpu.TALLY_ATOMIC, gpu.TALLY_DETERMINISTIC):
    d = driver.Driver(deck.write("/tmp/h.xml"), n_groups=30, device=0, validate=True, tally_mode=mode)
    v = d.gpu_context()
    r = d.cycle()
    post = v.download(gpu.LIST_WORK)
    E0 = post["E0"]; E = post["E"]; desc = post["descriptor"]
    n_new = r["gpu"]["n_new"]
    ex_in_new = math.fsum(E0[:n_new]); ex_in_cen = math.fsum(E0[n_new:])
    ex_abs = math.fsum(d.array("abs_E")); ex_exit = math.fsum(E[desc == 0]); ex_cen = math.fsum(E[desc == 2])
    tot = ex_in_new + ex_in_cen
    print("mode", mode, "rad_cons", r["rad_conservation"], "rel", r["rad_conservation"] / tot)
    print(" exact balance (fsum):", (ex_abs + ex_exit + ex_cen) - tot, "rel", ((ex_abs + ex_exit + ex_cen) - tot) / tot)
    print(" absorbed_E  reported-exact", r["absorbed_E"] - ex_abs)
    print(" exit_E      reported-exact", r["exit_E"] - ex_exit)
    print(" post_census reported-exact", r["post_census_E"] - ex_cen)
    print(" pre_census  reported-exact", r["pre_census_E"] - ex_in_cen)
    print(" emission+source reported-exact", r["emission_E"] + r["source_E"] - ex_in_new)
    d.close()
