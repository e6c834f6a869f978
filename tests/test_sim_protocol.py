"""The CPU model of the planned cluster / DSMEM Jacobi hand-over (tools/sim_jacobi_dsmem.py, mirrors the control flow of
jacobi_cluster_kernel in csrc/tail.cu): every pair of blocks meets once per sweep in its latest version, no slot is
overwritten or rotated while a neighbour reads it, no mbarrier ever has two outstanding phases, every block ends up home."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir, "tools"))


@pytest.mark.parametrize("NP,CS", [(2, 2), (3, 4), (7, 4), (9, 8), (16, 4), (63, 8)])
def test_handover_protocol_model(NP, CS):
    import sim_jacobi_dsmem as m
    for seed in range(4):
        m.Sim(NP, CS, 2, seed).run()
