"""CPU tests of the C++ host layer (no GPU, no compute calls on the device library).

* both shared libraries load and export every entry point their headers declare;
* Input reproduces the reference's parse rules (reference src/test/test_input.cc:83-95 and src/input.h:176-191,
  :488-490) and rejects malformed decks;
* Mesh reproduces the reference's mesh (reference src/test/test_mesh.cc: cell counts and region IDs) and, bit for bit,
  the oracle's face coordinates and per-cycle host quantities (f, op_a, E_emission, E_source, E_census, global source
  energy) -- the inputs the device consumes;
* time stepping equals the oracle's dt / next_dt sequence.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from branson_b200 import decks, driver, gpu
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_\w+)\s*\(", txt)))


def test_gpu_library_exports_every_declared_symbol():
    names = _declared("branson_gpu.h", "bgpu")
    assert len(names) >= 20
    L = C.CDLL(gpu.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/branson_gpu.h but not exported"
    assert set(gpu.EXPORTS) <= set(names)


def test_host_library_exports_every_declared_symbol():
    names = _declared("branson_host.h", "bhost")
    assert len(names) >= 10
    L = driver.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/branson_host.h but not exported"
    assert set(driver.EXPORTS) <= set(names)


def test_no_cpu_fallback_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    deck = decks.big_cube(n=4, photons=100, t_stop=0.001)
    with pytest.raises(driver.HostError, match="no CUDA device|CUDA error"):
        driver.Driver(deck.write(str(tmp_path / "d.xml")))
    d = driver.Driver(deck.write(str(tmp_path / "d.xml")), no_gpu=True)
    with pytest.raises(driver.HostError, match="no CPU transport"):
        d.cycle()


def test_input_parse_rules(tmp_path):
    # 1 rank + PARTICLE_PASS (or anything unknown) => REPLICATED; REPLICATED + HISTORY => batch_size 100000000
    deck = decks.hot_zone(photons=1234, t_stop=0.05, scale=10).with_(dd_transport_type="CELL_PASS", batch_size=777)
    d = driver.Driver(deck.write(str(tmp_path / "a.xml")), no_gpu=True)
    assert d.param("dd_mode") == 1  # Constants::REPLICATED
    assert d.param("particle_algorithm") == 0 and d.param("particle_storage") == 0
    assert d.param("batch_size") == 100000000
    assert d.param("n_user_photons") == 1234 and d.param("seed") == 14706
    assert [d.param(f"bc{i}") for i in range(6)] == [0.0] * 6
    assert d.param("n_cells") == deck.n_cells
    # EVENT keeps the user's batch size; SOA is recorded
    deck2 = deck.with_(dd_transport_type="REPLICATED", particle_algorithm="EVENT", particle_storage="SOA", batch_size=555)
    d2 = driver.Driver(deck2.write(str(tmp_path / "b.xml")), no_gpu=True)
    assert d2.param("batch_size") == 555 and d2.param("particle_algorithm") == 1 and d2.param("particle_storage") == 1
    # with more than one rank PARTICLE_PASS stays and is refused (the reference exits there too)
    with pytest.raises(driver.HostError, match="REPLICATED"):
        driver.Driver(deck.write(str(tmp_path / "a.xml")), no_gpu=True, rank=0, n_ranks=2)
    driver.Driver(deck.write(str(tmp_path / "a.xml")), no_gpu=True, rank=0, n_ranks=2, force_replicated=True)


def test_input_accepts_reference_deck_formatting(tmp_path):
    # leading blanks, '.65'-style numbers, comments and an XML declaration occur in the reference's decks
    xml = """<?xml version="1.0"?>
<!-- deck -->
<prototype>
  <common>
    <t_start> 0.0</t_start> <t_stop>.02</t_stop> <dt_start>  0.01 </dt_start> <t_mult>1.0</t_mult>
    <dt_max>1.0</dt_max> <photons> 500</photons> <seed>42</seed>
    <tilt>FALSE</tilt> <stratified_sampling>FALSE</stratified_sampling> <output_frequency>1</output_frequency>
    <dd_transport_type>REPLICATED</dd_transport_type>
  </common>
  <spatial>
    <x_division><x_start>0.0</x_start><x_end>.65</x_end><n_x_cells>3</n_x_cells></x_division>
    <y_division><y_start>0.0</y_start><y_end>1.0</y_end><n_y_cells>2</n_y_cells></y_division>
    <z_division><z_start>0.0</z_start><z_end>1.0</z_end><n_z_cells>1</n_z_cells></z_division>
    <region_map><x_div_ID>0</x_div_ID><y_div_ID>0</y_div_ID><z_div_ID>0</z_div_ID><region_ID>6</region_ID></region_map>
  </spatial>
  <boundary>
    <bc_right>VACUUM</bc_right><bc_left>SOURCE</bc_left><bc_up>REFLECT</bc_up><bc_down>REFLECT</bc_down>
    <bc_top>REFLECT</bc_top><bc_bottom>REFLECT</bc_bottom><T_source> 1.5</T_source>
  </boundary>
  <regions>
    <region><ID>6</ID><density>1.0</density><CV>2.0</CV><opacA>3.0</opacA><opacB>0.0</opacB><opacC>0.0</opacC>
      <opacS>0.0</opacS><initial_T_e>1.0</initial_T_e><initial_T_r>1.0</initial_T_r></region>
  </regions>
</prototype>
"""
    p = tmp_path / "deck.xml"
    p.write_text(xml)
    d = driver.Driver(str(p), no_gpu=True)
    assert d.param("n_cells") == 6 and d.param("n_user_photons") == 500 and d.param("seed") == 42
    assert d.param("t_stop") == 0.02 and d.param("dt") == 0.01 and d.param("T_source") == 1.5
    assert [d.param(f"bc{i}") for i in range(6)] == [3, 1, 0, 0, 0, 0]  # SOURCE VACUUM REFLECT...
    np.testing.assert_array_equal(d.array("x_faces"), [0.0, 0.0 + 1 * (0.65 / 3), 0.0 + 2 * (0.65 / 3), 0.0 + 3 * (0.65 / 3)])
    assert (d.array("T_s") > 0).sum() == 2  # only the two cells on the -x face carry the source temperature


@pytest.mark.parametrize("bad", ["no_common", "bad_bc", "region_map_count", "unknown_region", "garbage"])
def test_input_rejects_malformed_decks(tmp_path, bad):
    xml = decks.simple_three_region().to_xml()
    if bad == "no_common":
        xml = re.sub(r"<common>.*?</common>", "", xml, flags=re.S)
    elif bad == "bad_bc":
        xml = xml.replace("<bc_left>REFLECT</bc_left>", "<bc_left>MIRROR</bc_left>")
    elif bad == "region_map_count":
        xml = re.sub(r"<region_map>.*?</region_map>", "", xml, count=1, flags=re.S)
    elif bad == "unknown_region":
        xml = xml.replace("<region_ID>12</region_ID>", "<region_ID>99</region_ID>")
    else:
        xml = "<prototype><common></prototype>"
    p = tmp_path / "bad.xml"
    p.write_text(xml)
    with pytest.raises(driver.HostError):
        driver.Driver(str(p), no_gpu=True)


def test_mesh_like_reference_unit_test(tmp_path):
    # reference src/test/test_mesh.cc:41-113: simple_input.xml is 10 x 20 x 30 cells of region 6; the three-region deck
    # maps divisions to regions
    deck = decks.big_cube(n=4, photons=10, t_stop=0.001).with_(
        x_div=[(0.0, 1.0, 10)], y_div=[(0.0, 2.0, 20)], z_div=[(0.0, 3.0, 30)])
    d = driver.Driver(deck.write(str(tmp_path / "s.xml")), no_gpu=True)
    assert d.param("n_cells") == 10 * 20 * 30
    assert {d.param(f"region_of_cell:{i}") for i in (0, 17, 5999)} == {6.0}
    deck3 = decks.simple_three_region()
    d3 = driver.Driver(deck3.write(str(tmp_path / "t.xml")), no_gpu=True)
    sim = port.OracleSim(deck3)
    sim.cycle(keep_photons=False)
    want = sim.get("mesh/region")
    got = np.array([d3.param(f"region_of_cell:{i}") for i in range(deck3.n_cells)])
    np.testing.assert_array_equal(got, want)


CASES = {
    "three_region_g30": lambda: decks.simple_three_region(photons=3000, n_groups=30),
    "marshak": lambda: decks.marshak_wave(photons=3000, t_stop=0.03),
    "hot_zone_s10": lambda: decks.hot_zone(photons=5000, t_stop=0.03, scale=10),
    "hohlraum_s5": lambda: decks.hohlraum_single(photons=20000, t_stop=0.02, scale=5),
    "hohlraum_multi_s10": lambda: decks.hohlraum_multi(photons=8000, t_stop=0.003, scale=10),
    "big_cube_8": lambda: decks.big_cube(n=8, photons=6000, t_stop=0.002),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_mesh_and_first_cycle_host_quantities_equal_oracle_bitwise(name, tmp_path):
    deck = CASES[name]()
    d = driver.Driver(deck.write(str(tmp_path / "d.xml")), n_groups=deck.n_groups, no_gpu=True)
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=False)
    nx, ny, nz = deck.n_cells_xyz
    xf, yf, zf = gpu.faces_from_nodes(sim.get("mesh/nodes"), nx, ny, nz)
    for got, want in ((d.array("x_faces"), xf), (d.array("y_faces"), yf), (d.array("z_faces"), zf)):
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    gse = d.calculate_photon_energy()
    assert gse == sim.get("global_source_energy")[0]
    for k in ("f", "op_a", "op_s", "E_emission", "E_source", "E_census"):
        assert np.array_equal(d.array(k).view(np.uint64), sim.get(k).view(np.uint64)), k
    assert np.array_equal(d.array("T_e"), sim.get("T_e_pre"))


def test_time_stepping_equals_oracle(tmp_path):
    deck = decks.simple_three_region(photons=200)  # t_mult 1.5, dt_max 0.02: ramp, cap and end-of-run clip
    d = driver.Driver(deck.write(str(tmp_path / "d.xml")), no_gpu=True)
    sim = port.OracleSim(deck)
    n = 0
    while not sim.finished():
        assert not d.finished()
        sim.cycle(keep_photons=False)
        assert d.param("dt") == sim.get("dt")[0] and d.param("time") == sim.get("time")[0]
        assert d.param("next_dt") == sim.get("next_dt")[0]
        d.next_time_step()
        n += 1
    assert d.finished() and n == deck.n_cycles() and d.param("step") == n + 1


def test_aos_record_layout_matches_the_reference_photon():
    """gpu.aos_from_soa builds the reference's 120-byte Photon (src/photon.h:171-182: cell_ID u32, group u32, source_type
    u32, descriptors u8[4], pos f64[3], angle f64[3], E, E0, life_dx, RNG {ctr_lo, seed << 32, stream, 0}); the byte
    offsets are what bgpu_transport_photons_aos reads and writes (csrc/branson_gpu.cu k_aos_to_soa / k_soa_to_aos)."""
    import struct

    from branson_b200 import gpu
    soa = dict(cell=np.array([7, 9], np.uint32), group=np.array([3, 29], np.uint32),
               pos=np.array([.1, .2, .3, .4, .5, .6]), angle=np.array([1., 0., 0., 0., 0., -1.]),
               E=np.array([2.5, 3.5]), E0=np.array([4.5, 5.5]), life_dx=np.array([.7, .8]),
               ctr=np.array([11, 12], np.uint64), stream=np.array([10 ** 13 + 5, 10 ** 13 + 6], np.uint64))
    b = gpu.aos_from_soa(soa, seed=14706, source_type=np.array([2, 1])).tobytes()
    assert len(b) == 240
    rec = struct.unpack("<III4B3d3d3d4Q", b[120:])
    assert rec[:3] == (9, 29, 1) and rec[3] == gpu.PASS
    assert rec[7:10] == (.4, .5, .6) and rec[10:13] == (0., 0., -1.)
    assert rec[13:16] == (3.5, 5.5, .8)
    assert rec[16:] == (12, 14706 << 32, 10 ** 13 + 6, 0)
