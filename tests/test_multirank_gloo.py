"""world_size-2 gloo test (CPU): the sharded two-phase protocol of SURVEY.md §8e reproduces the single-process
result.  The compute stand-in per rank is the CPU oracle (test infrastructure); what is under test is the host
logic the CUDA path uses between the kernels: shard bounds, MIN all-reduce of {min_u,-max_u,min_tau,-max_tau},
SUM all-reduce of the additive partials, localization weight from the summed scatter."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "ma-lio_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import helpers as H
    import pyoracle as po
    from malio_b200 import dist as mdist, plugin, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth.make_case("mr", 4000, 40000, 3, 3, varied_map_cov=True)
    snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
    lo, hi = mdist.shard_bounds(case.pts.shape[0], rank, world)
    orc = po.Oracle(case.params)
    orc.set_map_snapshot(snap.nodes, snap.node_cov)
    orc.set_scan(case.pts[lo:hi], case.table, case.table_off, case.temporal_comp)
    # phase 1: local search + gates, local min/max
    orc.h_share_model(case.x_prop, True, 1)
    mm = orc.local_minmax()
    t = torch.tensor([mm[0], -mm[1], mm[2], -mm[3]], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    umin, umax, tmin, tmax = t[0].item(), -t[1].item(), t[2].item(), -t[3].item()
    # phase 2: weights with the global min/max, additive partials
    orc.set_minmax_override(True, umin, umax, tmin, tmax)
    orc.h_share_model(case.x_prop, False, 1)
    G, g, S = orc.partials()
    buf = torch.from_numpy(np.concatenate([G.ravel(), g, S, [orc.n_eff()]]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    c = orc.n_cols
    Gs = buf[:c * c].numpy().reshape(c, c); gs = buf[c * c:c * c + c].numpy(); Ss = buf[c * c + c:c * c + c + 6].numpy()
    w, sv = mdist.localization_weight(Ss, case.params)
    if rank == 0:
        np.savez(out, HTH=w * w * Gs, HTh=w * w * gs, n_eff=int(buf[-1].item()), w=w, mm=[umin, umax, tmin, tmax])
    dist.destroy_process_group()


def test_shard_bounds():
    from malio_b200 import dist as mdist
    for n in (0, 1, 7, 100000, 99999):
        for world in (1, 2, 3, 8):
            b = [mdist.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_protocol_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    import pyoracle as po
    from malio_b200 import plugin, synth
    out = str(tmp_path / "r0.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = np.load(out)
    case = synth.make_case("mr", 4000, 40000, 3, 3, varied_map_cov=True)
    snap = plugin.build_static_snapshot(case.map_xyz, case.map_normal_y)
    orc = H.make_oracle(case, snap)
    assert orc.h_share_model(case.x_prop, True, 2)
    HTH, HTh = orc.reduce()
    st = orc.stats()
    assert int(r["n_eff"]) == st.n_eff
    assert r["w"] == pytest.approx(st.loc_weight, rel=1e-10)
    np.testing.assert_allclose(r["mm"], [st.u_min, st.u_max, st.tau_min, st.tau_max], rtol=1e-13)
    np.testing.assert_allclose(r["HTH"], HTH, rtol=1e-9, atol=1e-9 * np.abs(HTH).max())
    np.testing.assert_allclose(r["HTh"], HTh, rtol=1e-9, atol=1e-9 * np.abs(HTh).max())
