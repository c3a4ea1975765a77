"""GPU parity: the CUDA hot path (through the C ABI, include/branson_gpu.h) against the oracle on the same inputs.

Every cycle the oracle (oracle/imc_oracle.c, pinned bit-for-bit to the unmodified reference) supplies the host-side
quantities of that cycle (f, op_a, op_s, per-rank E_emission / E_source / E_census, global source energy, dt, next_dt);
the device then sources, transports and compacts the census itself, and is compared photon by photon:

  * bit-exact: photon counts, per-photon cell, group, RNG stream, RNG counter (= number of draws), descriptor, and
    the per-photon event counters (events, scatters, cell crossings, reflections); census size and census order;
  * relative 1e-9 (as north_star states; typically 1e-13 is observed): positions, angles, energies, life_dx, per-cell
    abs_E / track_E, census_E, exit_E, pre_census_E.
"""
import numpy as np
import pytest

from branson_b200 import decks, gpu
from oracle import port

pytestmark = pytest.mark.gpu

RTOL = 1e-9


def _close(got, want, scale=None, what=""):
    got = np.asarray(got, np.float64)
    want = np.asarray(want, np.float64)
    assert got.shape == want.shape, what
    s = np.max(np.abs(want)) if scale is None else scale
    err = np.abs(got - want)
    tol = RTOL * np.abs(want) + RTOL * 1e-3 * s
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} differ, worst abs err {err.max():.3e} (scale {s:.3e})"
    return float(err.max() / s) if s > 0 else 0.0


_LOOKUPS = {}


def _run_cycles(deck, n_ranks=1, algorithm=gpu.HISTORY, tally_mode=gpu.TALLY_ATOMIC, max_cycles=None, launch=None,
                event_tail=None, closed_form_walk=True, group_arrays=False, tuning=None, event_queue=None):
    sim = port.OracleSim(deck, n_ranks=n_ranks)
    ctxs = None
    cyc = 0
    worst = 0.0
    while not sim.finished() and (max_cycles is None or cyc < max_cycles):
        cyc += 1
        sim.cycle(keep_photons=True)
        if ctxs is None:
            nodes = sim.get("mesh/nodes")
            ctxs = [gpu.context_for_deck(deck, nodes, rank=r, n_ranks=n_ranks, device=0) for r in range(n_ranks)]
            for c in ctxs:
                c.enable_counters(True)
                if launch:
                    c.set_launch(**launch)
                if event_tail is not None:  # the HBM-pass form of the event variant (csrc/event.cuh)
                    c.set_event_mode(hbm_passes=1)
                    c.set_event_tail(event_tail)
                if event_queue is not None:  # the shared-memory queue form (csrc/pool.cuh): election thresholds
                    c.set_event_mode(hbm_passes=0, batch_scatter=event_queue[0], batch_refill=event_queue[1])
                c.set_group_walk(closed_form_walk)
                if tuning:
                    c.set_divergence(tuning.get("scatter_batch", 0), tuning.get("aggregate", -1))
                    c.set_tally_copies(tuning.get("tally_copies", 0))
                    if "kernel" in tuning:
                        c.set_kernel(tuning["kernel"])
        dt, next_dt, gse = sim.get("dt")[0], sim.get("next_dt")[0], sim.get("global_source_energy")[0]
        f, op_a, op_s = sim.get("f"), sim.get("op_a"), sim.get("op_s")
        tot_abs = np.zeros(deck.n_cells)
        tot_trk = np.zeros(deck.n_cells)
        life_scale = 299.792458 * dt
        for r, ctx in enumerate(ctxs):
            if group_arrays:  # the general multigroup entry point, fed with the reference's faux-multigroup arrays
                G = deck.n_groups
                ctx.set_cell_groups(f, np.repeat(op_a, G), np.repeat(op_s, G))
            else:
                ctx.set_cell_data(f, op_a, op_s)
            n_new, n_tot = ctx.source(cyc, dt, sim.get("E_emission", r), sim.get("E_source", r),
                                      sim.get("E_census", r) if cyc == 1 else None, gse)
            assert n_new == int(sim.get("n_new", r)[0])
            assert n_tot == int(sim.get("n_photons", r)[0])
            pre = ctx.download(gpu.LIST_WORK)
            for k in ("cell", "group", "ctr", "stream"):
                assert np.array_equal(pre[k], sim.get("pre/" + k, r)), f"cycle {cyc} rank {r} pre/{k}"
            _close(pre["pos"], sim.get("pre/pos", r), what="pre/pos")
            _close(pre["angle"], sim.get("pre/angle", r), scale=1.0, what="pre/angle")
            _close(pre["E"], sim.get("pre/E", r), what="pre/E")
            _close(pre["E0"], sim.get("pre/E0", r), what="pre/E0")
            _close(pre["life_dx"], sim.get("pre/life_dx", r), scale=life_scale, what="pre/life_dx")

            ctx.transport(next_dt, algorithm, tally_mode)
            post = ctx.download(gpu.LIST_WORK, counters=True)
            for k in ("cell", "group", "ctr", "descriptor", "counters"):
                assert np.array_equal(post[k], sim.get("post/" + k, r)), f"cycle {cyc} rank {r} post/{k}"
            desc = post["descriptor"]
            # life_dx of census photons is reset by post-processing in the oracle dump and on the device alike
            _close(post["pos"], sim.get("post/pos", r), what="post/pos")
            _close(post["angle"], sim.get("post/angle", r), scale=1.0, what="post/angle")
            worst = max(worst, _close(post["E"], sim.get("post/E", r), scale=np.max(sim.get("pre/E0", r)), what="post/E"))
            not_census = desc != gpu.CENSUS
            _close(post["life_dx"][not_census], sim.get("post/life_dx", r)[not_census], scale=life_scale,
                   what="post/life_dx")

            a, t, st = ctx.tallies()
            worst = max(worst, _close(a, sim.get("rank_abs_E", r), what="rank_abs_E"))
            worst = max(worst, _close(t, sim.get("rank_track_E", r), what="rank_track_E"))
            tot_abs += a
            tot_trk += t
            assert st["n_census"] == int(sim.get("n_census", r)[0])
            assert st["n_transported"] == n_tot
            assert st["n_events"] == int(sim.get("post/counters", r)[0::4].astype(np.uint64).sum())
            assert st["n_scatters"] == int(sim.get("post/counters", r)[1::4].astype(np.uint64).sum())
            assert st["n_crossings"] == int(sim.get("post/counters", r)[2::4].astype(np.uint64).sum())
            assert st["n_reflections"] == int(sim.get("post/counters", r)[3::4].astype(np.uint64).sum())
            # (sigma_a, sigma_s) pairs of the algorithmic-bytes accounting: min(G, events) per visit
            visits = n_tot + st["n_crossings"]
            if deck.n_groups == 1:
                assert st["n_group_lookups"] == visits
            else:
                assert visits <= st["n_group_lookups"] <= min(st["n_events"], deck.n_groups * visits)
            # ... and a property of the workload: the same for every algorithm / tally mode / launch geometry that has
            # transported this deck in this session
            seen = _LOOKUPS.setdefault((deck.name, deck.n_groups, deck.photons, n_ranks, cyc, r), st["n_group_lookups"])
            assert st["n_group_lookups"] == seen
            assert st["n_killed"] == int((desc == gpu.KILLED).sum())
            assert st["n_exit"] == int((desc == gpu.EXIT).sum())
            e_scale = abs(gse)
            _close([st["census_E"]], sim.get("post_census_E", r), scale=e_scale, what="census_E")
            _close([st["exit_E"]], sim.get("exit_E", r), scale=e_scale, what="exit_E")
            _close([st["pre_census_E"]], sim.get("pre_census_E", r), scale=e_scale, what="pre_census_E")
            # per-rank radiation balance: what went in == what came out (IMC_State::print_conservation)
            e_in = st["pre_census_E"] + pre["E0"][:n_new].sum()
            e_out = a.sum() + st["census_E"] + st["exit_E"]
            assert abs(e_in - e_out) <= 1e-12 * e_in, f"radiation balance {e_in - e_out:.3e} of {e_in:.3e}"

            cen = ctx.download(gpu.LIST_CENSUS)
            sel = desc == gpu.CENSUS
            assert np.array_equal(cen["stream"], sim.get("pre/stream", r)[sel])
            assert np.array_equal(cen["cell"], sim.get("post/cell", r)[sel])
            assert np.array_equal(cen["ctr"], sim.get("post/ctr", r)[sel])
            assert np.all(cen["life_dx"] == 299.792458 * next_dt)
        _close(tot_abs, sim.get("abs_E"), what="abs_E")
        _close(tot_trk, sim.get("track_E"), what="track_E")
    for c in ctxs or []:
        c.close()
    return cyc, worst


def test_rng_known_answers_on_device():
    # same vectors as tests/test_oracle_golden.py (Random123 KATs; values printed by the reference build)
    assert gpu.threefry([0, 0], [0, 0]) == (0xc2b6e3a8c2c69865, 0x6f81ed42f350084d)
    m = 2 ** 64 - 1
    assert gpu.threefry([m, m], [m, m]) == (0xe02cb7c4d95d277a, 0xd06633d0893b8b68)
    assert gpu.threefry([0x243f6a8885a308d3, 0x13198a2e03707344], [0xa4093822299f31d0, 0x082efa98ec4e6c89]) == (
        0x263c7d30bb0f0af1, 0x56be8361d3311526)
    np.testing.assert_array_equal(gpu.rng_draws(777, 1, 4), [0.50060536157985436, 0.29511474933665161,
                                                             0.92517102853492827, 0.55080394150283085])
    np.testing.assert_array_equal(gpu.rng_draws(14706, 0, 4), [0.10704214220457897, 0.37725928982588541,
                                                               0.96811350044478395, 0.7575540980655614])
    np.testing.assert_array_equal(gpu.rng_draws(1234, 7, 2000), port.rng_draws(1234, 7, 2000))


CASES = {
    # name: (deck factory, n_ranks)
    "three_region_g1": (lambda: decks.simple_three_region(photons=20000, n_groups=1), 1),
    "three_region_g30_r2": (lambda: decks.simple_three_region(photons=20000, n_groups=30), 2),
    "marshak": (lambda: decks.marshak_wave(photons=20000, t_stop=0.05), 1),
    "hot_zone_s10": (lambda: decks.hot_zone(photons=30000, t_stop=0.04, scale=10), 1),
    "hohlraum_s5_g30": (lambda: decks.hohlraum_single(photons=60000, t_stop=0.03, scale=5), 1),
    "hohlraum_multi_s10_g30_r4": (lambda: decks.hohlraum_multi(photons=40000, t_stop=0.003, scale=10), 4),
    "big_cube_16": (lambda: decks.big_cube(n=16, photons=30000, t_stop=0.003), 1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_history_atomic_matches_oracle(name):
    mk, n_ranks = CASES[name]
    cyc, worst = _run_cycles(mk(), n_ranks=n_ranks)
    assert cyc >= 2
    assert worst < 1e-10


def test_full_mesh_hohlraum_matches_oracle_photon_by_photon():
    """The BASELINE mesh itself (65 x 65 x 140 cells, 30 groups, all 54 region blocks, reflecting and vacuum faces) at a
    photon count the oracle finishes in half a minute: 3e5 user photons, two cycles -- the streaming first cycle
    (~35 crossings per history) and a scattering one (~180 events per history, ~5e7 events), every per-photon integer
    bit-exact.  (The 1e7-photon runs of tests/test_gpu_fullsize.py check the same kernel through conservation laws.)"""
    cyc, worst = _run_cycles(decks.hohlraum_single(photons=300_000, t_stop=0.02), max_cycles=2)
    assert cyc == 2 and worst < 1e-10


@pytest.mark.parametrize("name", ["three_region_g30_r2", "marshak", "hohlraum_s5_g30"])
def test_history_deterministic_matches_oracle(name):
    mk, n_ranks = CASES[name]
    _run_cycles(mk(), n_ranks=n_ranks, tally_mode=gpu.TALLY_DETERMINISTIC)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("event_queue", [(24, 16), (1, 1), (32, 32), (7, 29)])
def test_event_queue_variant_matches_oracle(name, event_queue):
    """BGPU_EVENT, default form (csrc/pool.cuh: event queues in shared memory, two photon slots per lane, one event type
    per warp trip): the regrouping must not change a bit of any photon -- every per-photon integer against the oracle,
    for the default election thresholds, for 'serve every event at once' (1, 1), for 'wait for a full warp' (32, 32) and
    for an odd pair."""
    mk, n_ranks = CASES[name]
    _run_cycles(mk(), n_ranks=n_ranks, algorithm=gpu.EVENT, event_queue=event_queue, max_cycles=3)


def test_event_queue_variant_small_chunks_and_tally_copies():
    _run_cycles(decks.hot_zone(photons=20000, t_stop=0.02, scale=10), algorithm=gpu.EVENT,
                launch=dict(blocks_per_sm=1, chunk=7), tuning=dict(tally_copies=7))
    _run_cycles(decks.marshak_wave(photons=20000, t_stop=0.03), algorithm=gpu.EVENT, tuning=dict(tally_copies=64))


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("event_tail", [1, 2000])
def test_event_variant_matches_oracle(name, event_tail):
    """The event-based variant must reproduce the HISTORY results photon by photon (SURVEY 8a note N5): lockstep passes
    all the way down (tail 1) and with the history kernel finishing the last 2000 histories (RESUME path)."""
    mk, n_ranks = CASES[name]
    _run_cycles(mk(), n_ranks=n_ranks, algorithm=gpu.EVENT, event_tail=event_tail)


@pytest.mark.parametrize("name", ["three_region_g30_r2", "hohlraum_s5_g30"])
@pytest.mark.parametrize("group_arrays", [False, True])
def test_sequential_group_walk_matches_oracle(name, group_arrays):
    """sample_emission_group walked through the group array exactly like the reference (the default closed form is a
    provably equivalent shortcut for cells whose groups are all equal; both must give the oracle's integers)."""
    mk, n_ranks = CASES[name]
    _run_cycles(mk(), n_ranks=n_ranks, closed_form_walk=False, group_arrays=group_arrays)
    _run_cycles(mk(), n_ranks=n_ranks, closed_form_walk=True, group_arrays=group_arrays, max_cycles=2)


def test_deterministic_mode_is_bitwise_reproducible():
    deck = decks.hohlraum_single(photons=40000, t_stop=0.02, scale=5)
    outs = []
    for chunk in (128, 32):
        sim = port.OracleSim(deck)
        sim.cycle(keep_photons=False)
        ctx = gpu.context_for_deck(deck, sim.get("mesh/nodes"), device=0)
        ctx.set_launch(chunk=chunk)
        ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
        ctx.source(1, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"), sim.get("E_census"),
                   sim.get("global_source_energy")[0])
        ctx.transport(sim.get("next_dt")[0], gpu.HISTORY, gpu.TALLY_DETERMINISTIC)
        a, t, st = ctx.tallies()
        outs.append((a, t, st["census_E"], st["exit_E"]))
        ctx.close()
    assert np.array_equal(outs[0][0].view(np.uint64), outs[1][0].view(np.uint64))
    assert np.array_equal(outs[0][1].view(np.uint64), outs[1][1].view(np.uint64))
    assert outs[0][2:] == outs[1][2:]


def test_small_chunks_and_few_blocks_give_identical_photons():
    # work distribution must not change any per-photon result (SURVEY 8a note N5)
    _run_cycles(decks.hot_zone(photons=20000, t_stop=0.02, scale=10), launch=dict(blocks_per_sm=1, chunk=7))


@pytest.mark.parametrize("name", ["three_region_g30_r2", "marshak", "hot_zone_s10"])
@pytest.mark.parametrize("tuning", [
    dict(scatter_batch=1, aggregate=0, tally_copies=1),    # the plain loop: sample at once, one atomic pair per lane
    dict(scatter_batch=32, aggregate=1, tally_copies=1),   # scatters wait for a full warp; warp-aggregated deposits
    dict(scatter_batch=5, aggregate=1, tally_copies=7),    # odd sizes; replicated tallies folded after the launch
    dict(scatter_batch=12, aggregate=0, tally_copies=64),
    dict(kernel=gpu.KERNEL_QUEUES, tally_copies=7),        # BGPU_HISTORY served by the event-queue kernel (csrc/pool.cuh)
    dict(kernel=gpu.KERNEL_HISTORY),                       # ... and pinned to the history kernel (no auto switch)
])
def test_divergence_and_contention_knobs_do_not_change_results(name, tuning):
    """Parking scatters, combining same-cell deposits inside a warp and replicating the tally array only regroup work
    and reorder the tally sums: every per-photon integer stays bit-exact against the oracle, tallies within 1e-9."""
    mk, n_ranks = CASES[name]
    _run_cycles(mk(), n_ranks=n_ranks, tuning=tuning, max_cycles=3)


def _aos_from(pre, seed):
    return gpu.aos_from_soa(pre, seed, source_type=pre["source_type"]).copy()


@pytest.mark.parametrize("tally_mode", [gpu.TALLY_ATOMIC, gpu.TALLY_DETERMINISTIC])
def test_aos_drop_in_for_gpu_transport_photons(tally_mode):
    """bgpu_transport_photons_aos takes the reference's own 120-byte Photon records and Cell_Tally array
    (src/history_based_transport.h:348-413) and must update them like transport_photon does."""
    deck = decks.simple_three_region(photons=15000, n_groups=30)
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=True)
    pre = {k: sim.get("pre/" + k) for k in ("cell", "group", "source_type", "pos", "angle", "E", "E0", "life_dx", "ctr",
                                             "stream")}
    aos = _aos_from(pre, deck.seed)
    ctx = gpu.context_for_deck(deck, sim.get("mesh/nodes"), device=0)
    ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
    start = np.random.default_rng(5).random((deck.n_cells, 2)) * 1e-6  # tallies are accumulated, not overwritten
    tal = start.copy()
    ctx.transport_photons_aos(aos, tal, tally_mode=tally_mode)
    rec = aos.view(np.uint64).reshape(-1, 15)
    assert np.array_equal((rec[:, 0] & np.uint64(0xffffffff)).astype(np.uint32), sim.get("post/cell"))
    assert np.array_equal((rec[:, 0] >> np.uint64(32)).astype(np.uint32), sim.get("post/group"))
    assert np.array_equal(((rec[:, 1] >> np.uint64(32)) & np.uint64(0xff)).astype(np.uint8), sim.get("post/descriptor"))
    assert np.array_equal((rec[:, 1] & np.uint64(0xffffffff)).astype(np.uint32), pre["source_type"])
    assert np.array_equal(rec[:, 11], sim.get("post/ctr"))
    assert np.array_equal(rec[:, 13], pre["stream"])
    _close(rec[:, 2:5].copy().view(np.float64).reshape(-1), sim.get("post/pos"), what="aos pos")
    _close(rec[:, 8].copy().view(np.float64), sim.get("post/E"), scale=pre["E0"].max(), what="aos E")
    _close(tal[:, 0] - start[:, 0], sim.get("rank_abs_E"), what="aos abs_E")
    _close(tal[:, 1] - start[:, 1], sim.get("rank_track_E"), what="aos track_E")
    # a photon with a foreign RNG seed word must be rejected loudly, not silently mis-sampled
    bad = _aos_from(pre, deck.seed + 1)
    with pytest.raises(gpu.GpuError):
        ctx.transport_photons_aos(bad, tal)
    ctx.close()


def test_empty_and_no_cell_data_errors():
    deck = decks.big_cube(n=4, photons=100, t_stop=0.001)
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=False)
    ctx = gpu.context_for_deck(deck, sim.get("mesh/nodes"), device=0)
    z = np.zeros(deck.n_cells)
    with pytest.raises(gpu.GpuError):
        ctx.transport(0.001)  # cell data never set
    ctx.set_cell_data(sim.get("f"), sim.get("op_a"), sim.get("op_s"))
    n_new, n_tot = ctx.source(2, 0.001, z, z, None, 1.0)  # nothing to source, empty census
    assert (n_new, n_tot) == (0, 0)
    ctx.transport(0.001)
    a, t, st = ctx.tallies()
    assert not a.any() and not t.any() and st["n_census"] == 0 and st["census_E"] == 0.0
    ctx.close()


def test_group_dependent_scattering_is_not_treated_as_gray():
    """Cells whose groups share sigma_a but not sigma_s are genuinely multigroup: the photon's group enters the physics
    through the physical-vs-effective test, so none of the uniform-group shortcuts (closed-form walk, no reload on a
    group change, lazily sampled groups) may apply.  The event-based variant never takes those shortcuts: HISTORY must
    give the same photons, and the group dependence must be visible against the gray run."""
    deck = decks.simple_three_region(photons=20000, n_groups=30)
    sim = port.OracleSim(deck)
    sim.cycle(keep_photons=False)
    G = deck.n_groups
    f, op_a, op_s = sim.get("f"), sim.get("op_a"), sim.get("op_s")
    abs_groups = np.repeat(op_a, G)
    sct_gray = np.repeat(op_s, G)
    sct_groups = sct_gray * np.tile(1.0 + (np.arange(G) % 3), deck.n_cells)
    out = {}
    for key, algorithm, sct in (("history", gpu.HISTORY, sct_groups), ("event", gpu.EVENT, sct_groups),
                                ("gray", gpu.HISTORY, sct_gray)):
        ctx = gpu.context_for_deck(deck, sim.get("mesh/nodes"), device=0)
        ctx.enable_counters(True)
        ctx.set_cell_groups(f, abs_groups, sct)
        ctx.source(1, sim.get("dt")[0], sim.get("E_emission"), sim.get("E_source"), sim.get("E_census"),
                   sim.get("global_source_energy")[0])
        ctx.transport(sim.get("next_dt")[0], algorithm, gpu.TALLY_ATOMIC)
        out[key] = (ctx.download(gpu.LIST_WORK, counters=True), ctx.tallies())
        ctx.close()
    h, e, g = out["history"][0], out["event"][0], out["gray"][0]
    for k in ("cell", "group", "ctr", "descriptor", "counters"):
        assert np.array_equal(h[k], e[k]), k
    assert np.array_equal(h["E"].view(np.uint64), e["E"].view(np.uint64))
    assert out["history"][1][2]["n_group_lookups"] == out["event"][1][2]["n_group_lookups"]
    assert not np.array_equal(h["ctr"], g["ctr"])  # the test has power: group-dependent sigma_s changes the histories
