"""GPU tier: a pass-shaped call stream (tetwild_b200/callstream.py: the call mix one pass of MeshRefinement.cpp:120-183 makes)
replayed call by call through the C ABI, re-batched by kind, and call by call on the CPU oracle: same decisions, same energies."""
import numpy as np
import pytest

from tetwild_b200 import callstream as cs

pytestmark = pytest.mark.gpu


def test_pass_stream_three_ways(ctx, oracle):
    s = cs.make_pass_stream(250, 250, 150, seed=5)
    counts, units = cs.mix(s)
    assert all(counts[k] > 20 for k in cs.KINDS)
    G = cs.GpuReplayer(ctx, s)
    t_a, a = G.call_by_call()
    t_b, b = G.batched()
    G.close()
    t_c, c = cs.CpuReplayer(oracle, s).call_by_call()
    bad_ab = [i for i, (k, _) in enumerate(s["calls"]) if not cs.same(k, a[i], b[i], tol=0.0 if k != "newton" else 1e-12)]
    bad_ac = [i for i, (k, _) in enumerate(s["calls"]) if not cs.same(k, a[i], c[i])]
    assert not bad_ab, "call-by-call vs batched differ at %s" % [(i, s["calls"][i][0]) for i in bad_ab[:5]]
    assert not bad_ac, "GPU vs oracle differ at %s" % [(i, s["calls"][i][0]) for i in bad_ac[:5]]
    outs = [bool(a[i]) for i, (k, _) in enumerate(s["calls"]) if k == "faces_out"]
    assert 0.1 < np.mean(outs) < 0.95


def test_trial_energy_equals_move_evaluate_undo(ctx, oracle):
    """twg_mesh_vertex_trial_energy = the smoother's "move v, getNewEnergy(conn_tets[v]), move back" (VertexSmoother.cpp:505-541)
    without touching the resident mesh"""
    import tetwild_b200 as tw
    from tetwild_b200 import synth
    V, T = synth.grid_tet_mesh(6, 6, 6)
    M = tw.TetMesh(ctx, V, T)
    off, adj = M.get_rings()
    rng = np.random.default_rng(1)
    vs = rng.choice(len(V), 300).astype(np.int32)
    P = V[vs] + rng.normal(0, 0.01, size=(300, 3))
    P[7] = V[vs[7]] + 0.5                      # far outside its ring: inverted tets, still a finite positive energy or MAX_ENERGY
    E = M.vertex_trial_energy(vs, P)
    for j in range(0, 300, 7):
        v = int(vs[j])
        ring = adj[int(off[v]):int(off[v + 1])]
        M.set_vertices([v], P[j:j + 1])
        e = M.ring_energy(ring, np.array([0, len(ring)], dtype=np.uint64))[0]
        M.set_vertices([v], V[v:v + 1])
        assert e == E[j]
        V2 = V.copy()
        V2[v] = P[j]
        eo = oracle.amips_ring_energy(V2, T, np.array([0, len(ring)], dtype=np.uint64), t_ids=ring)[0]
        assert abs(e - eo) <= 1e-9 * abs(eo) or (e >= 1e49 and eo >= 1e49)
    assert np.array_equal(M.get_vertices(), V)     # the mesh was never left modified
    M.close()
