"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own, so these fixtures are outputs of the reference
itself on small seeded inputs:
  * distance_golden.npz : Distance::compare bit patterns (DistanceL2 / DistanceInnerProduct) on random pairs
  * roar_<name>.npz     : base, queries, reference-built RoarGraph index (raw file words), SearchRoarGraph
                          ids / dists / cmps / hops at several L_pq, and compute_groundtruth (MKL) kNN output
  * bipartite_<name>.npz: (python tests/golden/make_golden.py bipartite) for the inputs of roar_<name>.npz: the bipartite
                          graph file the reference's test_only_build_bipartite CLI writes at -T 1 (raw words), and
                          SearchBipartiteGraph ids / dists / cmps / hops + the 10 entry points per query from
                          oracle/_ref/libref_roargraph_detrand.so (the reference compiled with a fixed
                          std::random_device word per query, oracle/shims/fixed_random_device.h; word = seed + query)
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Reference, run_ref_compute_groundtruth  # noqa: E402
from roargraph_b200 import formats as F, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DIST_FN = {"ip": "mips", "l2": "l2", "cosine": "cosine"}


def distance_golden(R):
    rng = np.random.default_rng(7)
    rec = {}
    for dim in (8, 16, 24, 40, 96, 104, 200, 512):
        a = rng.standard_normal((64, dim)).astype(np.float32)
        b = rng.standard_normal((64, dim)).astype(np.float32)
        rec[f"a{dim}"] = a
        rec[f"b{dim}"] = b
        rec[f"ip{dim}"] = np.array([R.distance("ip", x, y) for x, y in zip(a, b)], dtype=np.float32)
        rec[f"l2{dim}"] = np.array([R.distance("l2", x, y) for x, y in zip(a, b)], dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "distance_golden.npz"), **rec)


def roar_golden(R, name, nb, ntrain, ntest, dim, metric, M_sq, M_pjbp, L_pjpq, Ls, K_gt, seed):
    base, train, test = synth.make_numpy(nb, ntrain, ntest, dim, seed=seed, latent=16)
    with tempfile.TemporaryDirectory() as d:
        bf, tf, qf = f"{d}/base.fbin", f"{d}/train.fbin", f"{d}/test.fbin"
        F.write_fbin(bf, base)
        F.write_fbin(tf, train)
        F.write_fbin(qf, test)
        run_ref_compute_groundtruth(bf, tf, f"{d}/train.gt", M_sq, DIST_FN[metric])
        run_ref_compute_groundtruth(bf, qf, f"{d}/test.gt", K_gt, DIST_FN[metric])
        # single-threaded build: the reference's build has benign races under OpenMP
        R.build_roargraph(bf, f"{d}/train.gt", f"{d}/index.rg", metric, M_sq, M_pjbp, L_pjpq, 1)
        index_words = np.fromfile(f"{d}/index.rg", dtype=np.uint32)
        gt_ids, gt_dists = F.read_truthset(f"{d}/test.gt")
        tr_ids, _ = F.read_truthset(f"{d}/train.gt")
        pdim = (dim + 7) // 8 * 8
        h = R.index_load(bf, f"{d}/index.rg", metric, pdim, 1)
        q = F.pad_rows(test)
        extra = {}
        if metric == "cosine":
            # -Ofast vectorises the reference's normalize<T> (util.h:214-225), so its rounding is a property of the
            # compiled binary; keep the rows exactly as the reference's loader / CLI produced them.
            q = R.normalize_rows(q)
            extra = dict(base_norm=R.index_base(h, nb, pdim).copy(), queries_norm=q)
        rec = dict(base=base, test=test, train_knn_ids=tr_ids, index_words=index_words, gt_ids=gt_ids, gt_dists=gt_dists,
                   metric=np.array(metric), Ls=np.array(Ls, dtype=np.uint32), k=np.array(10, dtype=np.uint32),
                   params=np.array([M_sq, M_pjbp, L_pjpq], dtype=np.uint32), **extra)
        for L in Ls:
            r = R.search_roargraph(h, q, 10, L, 1)
            rec[f"ids_L{L}"] = r["ids"]
            rec[f"dists_L{L}"] = r["dists"]
            rec[f"cmps_L{L}"] = r["cmps"]
            rec[f"hops_L{L}"] = r["hops"]
        R.index_free(h)
    np.savez_compressed(os.path.join(OUT, f"roar_{name}.npz"), **rec)
    print(name, "index words", index_words.size, "avg cmps", {L: float(rec[f"cmps_L{L}"].mean()) for L in Ls})


def bipartite_golden(name, M_pjbp=16, cases=((10, 1), (64, 2), (300, 3)), k=5):
    import subprocess
    Rd = Reference(detrand=True)
    z = np.load(os.path.join(OUT, f"roar_{name}.npz"))
    metric = str(z["metric"])
    base, test, knn = z["base"], z["test"], z["train_knn_ids"]
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    with tempfile.TemporaryDirectory() as d:
        bf, tf, kf, gf = f"{d}/base.fbin", f"{d}/train.fbin", f"{d}/train.gt", f"{d}/bp.rg"
        F.write_fbin(bf, base)
        F.write_fbin(tf, np.zeros((knn.shape[0], base.shape[1]), np.float32))  # only its header is read on this path
        F.write_truthset(kf, knn, np.zeros(knn.shape, np.float32))
        subprocess.check_call([os.path.join(ref_dir, "ref_test_only_build_bipartite"), "--data_type", "float", "--dist", metric,
                               "--base_data_path", bf, "--sampled_query_data_path", tf, "--bipartite_index_save_path", gf,
                               "--projection_index_save_path", gf + ".unused", "--learn_base_nn_path", kf,
                               "--base_learn_nn_path", kf, "--M_pjbp", str(M_pjbp), "-T", "1"], stdout=subprocess.DEVNULL)
        words = np.fromfile(gf, dtype=np.uint32)
        pdim = (base.shape[1] + 7) // 8 * 8
        h = Rd.bipartite_load(bf, gf, metric, pdim, knn.shape[0], 1)
        q = F.pad_rows(test)
        if metric == "cosine":
            q = Rd.normalize_rows(q)
        rec = dict(graph_words=words, M_pjbp=np.array(M_pjbp, np.uint32), k=np.array(k, np.uint32),
                   cases=np.array(cases, dtype=np.uint32))
        for L, seed in cases:
            r = Rd.search_bipartite(h, q, k, L, 1, seed_base=seed)
            for key in ("ids", "dists", "cmps", "hops", "init_ids"):
                rec[f"{key}_L{L}"] = r[key]
        Rd.index_free(h)
    np.savez_compressed(os.path.join(OUT, f"bipartite_{name}.npz"), **rec)
    print("bipartite", name, "graph words", words.size, {L: float(rec[f"cmps_L{L}"].mean()) for L, _ in cases})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bipartite":  # adds the bipartite fixtures without touching the others
        for nm in ("ip_d40", "l2_d104", "cos_d64"):
            bipartite_golden(nm)
        sys.exit(0)
    R = Reference()
    distance_golden(R)
    roar_golden(R, "ip_d40", 4000, 4000, 64, 40, "ip", 32, 12, 48, (10, 32, 100), 20, 11)
    roar_golden(R, "l2_d104", 2000, 2000, 48, 104, "l2", 32, 12, 48, (10, 50), 20, 12)
    roar_golden(R, "cos_d64", 2000, 2000, 48, 64, "cosine", 32, 12, 48, (10, 50), 20, 13)
