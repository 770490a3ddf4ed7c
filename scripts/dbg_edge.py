import sys; sys.path.insert(0,'/root/repo')
import numpy as np, oracle
import tetwild_b200 as tw
ctx=tw.Context(0)
V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
P = np.random.default_rng(0).uniform(-0.5, 1.5, size=(5000, 3)); P[:100, 2] = 0.0
for F in (np.array([[0, 1, 2]]), np.array([[0, 1, 2], [0, 1, 3]]), np.array([[0, 1, 2], [0, 1, 3], [1, 2, 3]]), np.array([[0, 1, 1], [1, 2, 2], [2, 3, 3], [0, 1, 2]])):
    F=F.astype(np.uint32)
    S = tw.Surface(ctx, V, F); OS = oracle.Surface(V, F)
    for eps2 in (0.0, 1e-4, 1e-2, 0.3):
        a=S.points_out(P, eps2); b=OS.points_out(P, eps2); bb=OS.points_out(P,eps2,brute=True)
        print(len(F), eps2, "mismatch vs tree", (a!=b).sum(), "vs brute", (a!=bb).sum(), "tree vs brute", (b!=bb).sum(), a.mean())
    d=S.nearest(P)[2]; do=OS.nearest(P)[2]; db=OS.sqdist_brute(P)[0]
    print(len(F), "nearest mismatch", (d!=do).sum(), (d!=db).sum(), np.abs(d-do).max())
    i=np.nonzero(d!=do)[0][:3]; print(i, d[i], do[i], db[i])
