"""Generates tests/golden/*.npz: oracle chains on small seeded problems (inputs are regenerated from the
seed; the fixture stores the variate log and the per-iteration states).  The reference has no golden
vectors of its own (test/runtests.jl is empty) and cannot run here (no Julia), so these pin the ORACLE
against regressions and give the GPU tests a committed target that does not need the oracle .so.
Run:  python tests/golden/make_golden.py [extra | weighted]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from common import make_problem, oracle_chain  # noqa: E402

CASES = {
    "bayespr_rr": dict(n=300, p=150, seed=101, method=0, v=0.01, iters=25),
    "bayespr_regions": dict(n=257, p=130, seed=102, method=0, v=0.01, iters=25, region_off=[0, 40, 41, 100, 130]),
    "bayesb": dict(n=300, p=150, seed=103, method=1, v=0.05, pi=0.1, est_pi=True, iters=25),
    "bayesc_pi": dict(n=311, p=200, seed=104, method=2, v=0.05, pi=0.05, est_pi=True, iters=25),
}


EXTRA = {
    "bayesr": dict(n=280, p=120, seed=105, v=0.5, est_pi=True, iters=20, v_class=[0.0, 0.0001, 0.001, 0.01], pi=[0.8, 0.1, 0.07, 0.03]),
    "tuple2": dict(n=260, p=90, seed=106, iters=15, V=[[0.02, 0.005], [0.005, 0.03]], region_off=[0, 30, 90]),
}


def breeds(n, p, k, seed):
    probs = [make_problem(n, p, seed + 17 * b) for b in range(k)]
    y = probs[0]["y"].copy()
    for b in range(1, k):
        y += probs[b]["y"] - probs[b]["y"].mean()
    return probs, y


def main_extra():
    from oracle import oracle as O
    c = EXTRA["bayesr"]
    prob = make_problem(c["n"], c["p"], c["seed"])
    X, mean, mpm = O.center_codes(prob["codes"])
    R = O.BayesROracle(X, mpm, np.array(c["pi"]), np.array(c["v_class"]), v=c["v"], est_pi=c["est_pi"])
    ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2)
    out = {k: [] for k in ("chi2_e", "z_mu", "u", "z", "chi2_b", "dir_pi", "beta", "delta", "varBeta", "varE", "mu", "pi")}
    for _ in range(c["iters"]):
        log = ch.iteration(seed=c["seed"], chain=3)
        s = R.sweep(ch.e, ch.varE, it=ch.iter, seed=c["seed"], chain=3)
        out["chi2_e"].append(log["chi2_e"]); out["z_mu"].append(log["z_mu"])
        for k in ("u", "z", "chi2_b", "dir_pi"):
            out[k].append(np.array(s[k]))
        out["beta"].append(R.beta.copy()); out["delta"].append(R.delta.copy()); out["varBeta"].append(R.varBeta.copy())
        out["varE"].append(ch.varE); out["mu"].append(ch.mu); out["pi"].append(R.piHat.copy())
    np.savez_compressed(os.path.join(HERE, "bayesr.npz"), e_final=ch.e, **{k: np.array(v) for k, v in out.items()})
    print("bayesr varE", ch.varE, "classes", np.bincount(R.delta, minlength=5)[1:])
    c = EXTRA["tuple2"]
    probs, y = breeds(c["n"], c["p"], 2, c["seed"])
    Xk = [O.center_codes(pr["codes"])[0] for pr in probs]
    mb = O.MultiBreedOracle(Xk, np.array(c["V"]), region_off=np.array(c["region_off"], dtype=np.int64))
    ch = O.OracleChain(y, [], v_e=float(np.var(y)) / 2)
    out = {k: [] for k in ("chi2_e", "z_mu", "z", "iw_chi2", "iw_z", "beta", "varBeta", "varE", "mu")}
    for _ in range(c["iters"]):
        log = ch.iteration(seed=c["seed"], chain=3)
        s = mb.sweep(ch.e, ch.varE, it=ch.iter, seed=c["seed"], chain=3)
        out["chi2_e"].append(log["chi2_e"]); out["z_mu"].append(log["z_mu"])
        for k in ("z", "iw_chi2", "iw_z"):
            out[k].append(np.array(s[k]))
        out["beta"].append(mb.beta.copy()); out["varBeta"].append(mb.varBeta.copy()); out["varE"].append(ch.varE); out["mu"].append(ch.mu)
    np.savez_compressed(os.path.join(HERE, "tuple2.npz"), e_final=ch.e, **{k: np.array(v) for k, v in out.items()})
    print("tuple2 varE", ch.varE)


WEIGHTED = {
    "bayesc_weighted": dict(n=290, p=140, seed=107, method=2, v=0.05, pi=0.1, est_pi=True, iters=20),
    "bayesb_weighted": dict(n=270, p=110, seed=108, method=1, v=0.05, pi=0.2, est_pi=True, iters=20),
}


def weights_of(c):
    """Residual weights w = E.iVarStr of the weighted cases (regenerated from the seed, like the genotypes)."""
    return np.random.default_rng(c["seed"]).uniform(0.25, 4.0, c["n"])


def weighted_chain(c):
    from oracle import oracle as O
    prob = make_problem(c["n"], c["p"], c["seed"])
    w = weights_of(c)
    X, mean, _ = O.center_codes(prob["codes"])
    Mp, mpm = O.weighted_marker_arrays(X, w)                       # mme.jl:299-303
    S = O.MarkerSet(X=X, mpm=mpm, method=c["method"], v=c["v"], pi=c["pi"], est_pi=c["est_pi"])
    S.Mp = Mp
    return prob, w, O.OracleChain(prob["y"], [S], v_e=prob["var_y"] / 2, weights=w), S


def main_weighted():
    """E.str == "D" (SURVEY §8 f3): oracle chains with residual weights."""
    for name, c in WEIGHTED.items():
        prob, w, ch, S = weighted_chain(c)
        out = {k: [] for k in ("chi2_e", "z_mu", "u", "z", "chi2_b", "beta_pi", "beta", "delta", "varBeta", "varE", "mu", "pi")}
        for _ in range(c["iters"]):
            log = ch.iteration(seed=c["seed"], chain=3)
            s = log["sets"][0]
            for k in ("chi2_e", "z_mu"):
                out[k].append(log[k])
            for k in ("u", "z", "chi2_b", "beta_pi"):
                out[k].append(np.array(s[k]))
            out["beta"].append(S.beta.copy()); out["delta"].append(S.delta.copy()); out["varBeta"].append(S.varBeta.copy())
            out["varE"].append(ch.varE); out["mu"].append(ch.mu); out["pi"].append(S.piHat.copy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), e_final=ch.e, **{k: np.array(v) for k, v in out.items()})
        print(name, "varE", ch.varE, "nIn", int(S.delta.sum()))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "extra":
        return main_extra()
    if len(sys.argv) > 1 and sys.argv[1] == "weighted":
        return main_weighted()
    for name, c in CASES.items():
        prob = make_problem(c["n"], c["p"], c["seed"])
        ro = np.array(c["region_off"], dtype=np.int64) if "region_off" in c else None
        ch, S = oracle_chain(prob, c["method"], c["v"], pi=c.get("pi", 0.0), est_pi=c.get("est_pi", False), region_off=ro)
        out = {k: [] for k in ("chi2_e", "z_mu", "u", "z", "chi2_b", "beta_pi", "beta", "delta", "varBeta", "varE", "mu", "pi")}
        for _ in range(c["iters"]):
            log = ch.iteration(seed=c["seed"], chain=3)
            s = log["sets"][0]
            for k in ("chi2_e", "z_mu"):
                out[k].append(log[k])
            for k in ("u", "z", "chi2_b", "beta_pi"):
                out[k].append(np.array(s[k]))
            out["beta"].append(S.beta.copy()); out["delta"].append(S.delta.copy()); out["varBeta"].append(S.varBeta.copy())
            out["varE"].append(ch.varE); out["mu"].append(ch.mu); out["pi"].append(S.piHat.copy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), e_final=ch.e, **{k: np.array(v) for k, v in out.items()})
        print(name, "varE", ch.varE, "nIn", int(S.delta.sum()))


RC = {
    "bayesrc_pi": dict(n=240, p=80, seed=107, v=0.6, plus=False, iters=12, n_annot=3, v_class=[0.0, 0.001, 0.01, 0.1], pi=[0.7, 0.15, 0.1, 0.05]),
    "bayesrc_plus": dict(n=240, p=80, seed=108, v=0.6, plus=True, iters=12, n_annot=3, v_class=[0.0, 0.001, 0.01, 0.1], pi=[0.7, 0.15, 0.1, 0.05]),
}


def rc_annot(p, nA, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2, size=(p, nA)).astype(np.int32)
    a[a.sum(1) == 0, 0] = 1
    a[1, :] = np.arange(nA, 0, -1)
    return a


def main_rc():
    """BayesRCpi / BayesRCplus (functions.jl:291-419): variate logs and per-iteration states of the oracle (python make_golden.py rc)"""
    from oracle import oracle as O
    for name, c in RC.items():
        prob = make_problem(c["n"], c["p"], c["seed"])
        X, _, mpm = O.center_codes(prob["codes"])
        annot = rc_annot(c["p"], c["n_annot"], c["seed"])
        R = O.BayesRCOracle(X, mpm, np.array(c["pi"]), np.array(c["v_class"]), v=c["v"], annot=annot, est_pi=True, plus=c["plus"])
        ch = O.OracleChain(prob["y"], [], v_e=prob["var_y"] / 2)
        keys = ("u_annot", "dirp", "u", "z", "chi2_b", "dir_pi")
        out = {k: [] for k in ("chi2_e", "z_mu", "beta", "delta", "varBeta", "varE", "mu", "pi", "annot_cat") + keys}
        for _ in range(c["iters"]):
            log = ch.iteration(seed=c["seed"], chain=2)
            s = R.sweep(ch.e, ch.varE, it=ch.iter, seed=c["seed"], chain=2)
            out["chi2_e"].append(log["chi2_e"]); out["z_mu"].append(log["z_mu"])
            for k in keys:
                out[k].append(np.array(s[k]))
            out["beta"].append(R.beta.copy()); out["delta"].append(R.delta.copy()); out["varBeta"].append(R.varBeta.copy())
            out["varE"].append(ch.varE); out["mu"].append(ch.mu); out["pi"].append(R.piHat.copy()); out["annot_cat"].append(R.annot_cat.copy())
        np.savez_compressed(os.path.join(HERE, name + ".npz"), e_final=ch.e, annot_prob_final=R.annot_prob, **{k: np.array(v) for k, v in out.items()})
        print(name, "varE", ch.varE, "classes", np.bincount(R.delta, minlength=5)[1:], "varBeta", R.varBeta)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rc":
        main_rc()
    else:
        main()
