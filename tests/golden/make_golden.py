"""Generates tests/golden/*.npz from the oracle (which drives the REAL LAPACK dlaqps/dgemm/dtrsm of
scipy's OpenBLAS).  The Julia reference cannot run in this image and ships no golden vectors of its own
(SURVEY.md section 4), so these fixtures pin (a) the oracle against silent drift and (b) the CUDA path
against fixed inputs on the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, os.pardir, os.pardir, "oracle"))
import lra_oracle as o  # noqa: E402


def qrcp_case(name, B0, **kw):
    B = B0.copy(order="F")
    tr = o.QRCPTrace()
    p, tau, k = o.geqp3_adap(B, o.LRAOptions(**kw), o.dlaqps_real, tr)
    ns = tr.steps
    np.savez_compressed(os.path.join(HERE, name + ".npz"), B0=B0, p=p, k=k, kb=np.array(tr.kb), steps=ns,
                        R=np.triu(B[:ns, :]), tau=tau[:ns], opts=np.array(sorted(kw.items()), dtype=object))


def id_case(name, A, seed, trans="n", **kw):
    rin = o.RandomInputs(seed)
    V = o.idfact(A, o.LRAOptions(**kw), rin, trans)
    d = {"A": A, "sk": V.sk, "rd": V.rd, "T": V.T, "rounds": np.array(V.rounds), "trans": trans,
         "opts": np.array(sorted(kw.items()), dtype=object)}
    for t, r in enumerate(rin.drawn):
        for key, val in r.items():
            if key not in ("kind", "round", "order"):
                d[f"rand{t}_{key}"] = val
    d["n_rand"] = len(rin.drawn)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def _rand_fields(rin, d):
    for t, r in enumerate(rin.drawn):
        for key, val in r.items():
            if key not in ("kind", "round", "order"):
                d[f"rand{t}_{key}"] = val
    d["n_rand"] = len(rin.drawn)


def psvd_case(name, A, seed, **kw):
    """psvdfact + pqrfact of the same matrix on the same draws; factors stored with the sign convention
    'largest-magnitude entry of every left vector positive' so that any correct implementation can be compared."""
    rin = o.RandomInputs(seed)
    F = o.psvdfact(A, o.LRAOptions(**kw), rin)
    sgn = np.sign(F.U[np.argmax(np.abs(F.U), axis=0), np.arange(F.U.shape[1])])
    d = {"A": A, "S": F.S, "US": F.U * sgn * F.S, "SVt": (F.Vt * sgn[:, None]) * F.S[:, None], "k_id": F.k_id,
         "opts": np.array(sorted(kw.items()), dtype=object)}
    _rand_fields(rin, d)
    trans = "n" if A.shape[0] >= A.shape[1] else "c"          # the side psvdfact factors (src/psvd.jl:242,256): same draws
    rin2 = o.RandomInputs(seed)
    Q = o.pqrfact(A, o.LRAOptions(**kw), rin2, trans)
    dq = np.sign(np.diag(Q.R[:, :Q.Q.shape[1]]))
    d.update({"p": Q.p, "Q": Q.Q * dq, "R": Q.R * dq[:, None], "trans": trans})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def new_cases_r01f():
    """Fixtures added late in round 1 (kept apart so that regenerating them never rewrites the older files)."""
    A = o.decaying_matrix(120, 90, 36, 11.0, 36, seed=7)
    psvd_case("psvd_decay_120x90", A, seed=6, rtol=1e-10)
    psvd_case("psvd_decay_90x120", np.asfortranarray(A.T), seed=8, rtol=1e-9)


def main():
    rng = np.random.default_rng(2024)
    A = o.decaying_matrix(96, 80, 40, 13.0, 40, seed=1)
    qrcp_case("qrcp_decay_40x80", o.sketch_randn(A, np.asfortranarray(rng.standard_normal((40, 96)))), rtol=1e-12)
    qrcp_case("qrcp_gauss_24x48_rank10", np.asfortranarray(rng.standard_normal((24, 48))), rank=10)
    H = o.matrixlib_hilb(64)
    qrcp_case("qrcp_hilb_40x64", o.sketch_randn(H, np.asfortranarray(rng.standard_normal((40, 64)))), rtol=1e-10)
    id_case("id_decay_96x80", A, seed=3, rtol=1e-11)
    id_case("id_decay_96x80_c", A, seed=4, trans="c", rtol=1e-11)
    x, y = np.sort(rng.random(64)), 1.02 + np.sort(rng.random(64))
    id_case("id_cauchy_64", o.matrixlib_cauchy(x, y), seed=5, rtol=1e-10)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "r01f":
        new_cases_r01f()
    else:
        main()
        new_cases_r01f()
