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
    main()
