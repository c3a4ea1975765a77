"""Fingerprints of the reference's own input decks, as parsed by the C++ host layer (TEST INFRASTRUCTURE).

    python oracle/gen_deck_fixtures.py            # needs /root/reference/inputs; writes tests/golden/reference_decks.json

branson_b200/decks.py generates the BASELINE decks from the reference decks' numbers (the GPU box has no /root/reference).
This script pins that: every reference XML is read by `Input` / `Mesh` (no_gpu) and reduced to a fingerprint -- the parsed
scalars plus SHA-256 digests of the mesh faces, the initial temperatures and the first cycle's host quantities (f, op_a,
E_emission, E_census, E_source: they depend on every region property and on the region map).  tests/test_decks.py
demands the same fingerprints from the generated decks.  big_cube.xml (800^3 cells = 5.1e8) is fingerprinted from its
XML text only (decks.big_cube is the scaled 200^3 problem; its per-cell numbers are the deck's).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import xml.etree.ElementTree as ET

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_INPUTS = "/root/reference/inputs"
OUT = os.path.join(ROOT, "tests", "golden", "reference_decks.json")
# reference deck -> (BRANSON_N_GROUPS the BASELINE config builds it with, force REPLICATED)
DECKS = {
    "marshak_wave_replicated.xml": (1, False),
    "hot_zone_input.xml": (1, False),
    "3D_hohlraum_single_node.xml": (30, False),
    "3D_hohlraum_multi_node.xml": (30, True),
}
SCALARS = ["n_cells", "nx", "ny", "nz", "n_user_photons", "seed", "dd_mode", "particle_algorithm", "particle_storage",
           "batch_size", "use_gpu_transporter", "use_comb", "t_start", "t_stop", "dt", "t_mult", "dt_max", "T_source",
           "n_regions", "bc0", "bc1", "bc2", "bc3", "bc4", "bc5"]
ARRAYS = ["x_faces", "y_faces", "z_faces", "T_e", "T_s", "f", "op_a", "op_s", "E_emission", "E_census", "E_source"]


def fingerprint(xml_path: str, n_groups: int, force_replicated: bool) -> dict:
    from branson_b200 import driver
    d = driver.Driver(xml_path, n_groups=n_groups, no_gpu=True, force_replicated=force_replicated)
    try:
        fp = {"scalars": {k: d.param(k) for k in SCALARS}}
        fp["global_source_energy"] = d.calculate_photon_energy()
        fp["arrays"] = {k: hashlib.sha256(d.array(k).tobytes()).hexdigest() for k in ARRAYS}
    finally:
        d.close()
    return fp


def big_cube_text_fingerprint(xml_path: str) -> dict:
    """the numbers decks.big_cube takes from inputs/big_cube.xml (everything except the cell counts it scales)"""
    root = ET.parse(xml_path).getroot()
    txt = lambda path: root.find(path).text.strip()  # noqa: E731
    sp = root.find("spatial")
    reg = root.find("regions/region")
    return {"dt_start": float(txt("common/dt_start")), "t_mult": float(txt("common/t_mult")),
            "dt_max": float(txt("common/dt_max")), "seed": int(txt("common/seed")),
            "cell_size": (float(sp.find("x_division/x_end").text) - float(sp.find("x_division/x_start").text))
            / int(sp.find("x_division/n_x_cells").text),
            "bc": [root.find(f"boundary/bc_{s}").text.strip() for s in ("left", "right", "down", "up", "bottom", "top")],
            "region": {k: float(reg.find(k).text) for k in ("density", "CV", "opacA", "opacB", "opacC", "opacS",
                                                            "initial_T_e", "initial_T_r")}}


def main():
    out = {}
    for name, (g, force) in DECKS.items():
        out[name] = fingerprint(os.path.join(REF_INPUTS, name), g, force)
        print(name, out[name]["scalars"]["n_cells"], out[name]["global_source_energy"])
    out["big_cube.xml"] = big_cube_text_fingerprint(os.path.join(REF_INPUTS, "big_cube.xml"))
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
