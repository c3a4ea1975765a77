"""Run the UNMODIFIED reference (oracle/_ref/ref_harness_g*) and read its binary dumps.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DT = {0: np.float64, 1: np.uint32, 2: np.uint64, 3: np.uint8}


def harness_path(n_groups: int, dropin: bool = False) -> str:
    """dropin: the reference with oracle/dropin.patch applied (its gpu_transport_photons calls the product library)"""
    return os.path.join(_HERE, "_ref", f"ref_harness_{'dropin_' if dropin else ''}g{n_groups}")


def stock_binary_path(n_groups: int, dropin: bool = False) -> str:
    return os.path.join(_HERE, "_ref", f"branson_ref_{'dropin_' if dropin else ''}g{n_groups}")


def have_reference(n_groups: int) -> bool:
    return os.path.exists(harness_path(n_groups))


def read_dump(path: str) -> dict:
    out = {}
    with open(path, "rb") as fh:
        data = fh.read()
    off = 0
    while off < len(data):
        (nl,) = struct.unpack_from("<I", data, off)
        off += 4
        name = data[off:off + nl].decode()
        off += nl
        dt = data[off]
        off += 1
        (cnt,) = struct.unpack_from("<Q", data, off)
        off += 8
        t = np.dtype(_DT[dt])
        out[name] = np.frombuffer(data, dtype=t, count=cnt, offset=off).copy()
        off += cnt * t.itemsize
    return out


def run_reference(deck, n_ranks: int = 1, max_cycles: int | None = None, photon_limit: int | None = None,
                  workdir: str | None = None, timeout: float = 3600.0, comb_max: int | None = None,
                  comb_stream: int = 0, dump_level: int = 0, dropin: bool = False):
    """Returns (list of per-rank dump dicts, stdout of rank 0).  dump_level 1: per-photon integers of the
    post-transport list, abs_E / track_E / T_e / T_r and the scalars only (full-size runs)."""
    exe = harness_path(deck.n_groups, dropin)
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} (build with `make -C oracle ref{'_dropin' if dropin else ''}` where "
                                "/root/reference exists)")
    tmp = workdir or tempfile.mkdtemp(prefix="branson_ref_")
    xml = deck.write(os.path.join(tmp, deck.name + ".xml"))
    prefix = os.path.join(tmp, deck.name)
    env = dict(os.environ)
    env["BRANSON_SHIM_NRANKS"] = str(n_ranks)
    if dump_level:
        env["REF_DUMP_LEVEL"] = str(int(dump_level))
    if comb_max is not None:  # comb_photons on the final census (dumped under comb/...)
        env["REF_COMB_MAX"] = str(int(comb_max))
        env["REF_COMB_STREAM"] = str(int(comb_stream))
    args = [exe, xml, prefix, str(max_cycles if max_cycles is not None else 2 ** 31 - 1)]
    if photon_limit is not None:
        args.append(str(photon_limit))
    res = subprocess.run(args, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"reference harness failed ({res.returncode}):\n{res.stdout[-4000:]}")
    dumps = [read_dump(f"{prefix}.rank{r}.bin") for r in range(n_ranks)]
    return dumps, res.stdout
