import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import helpers as H
    from foldcomp_b200 import abi

    z = np.load(os.path.join(HERE, "golden", "golden.npz"))
    batch = abi.HostChainBatch(
        res_off=z["res_off"], atom_off=z["atom_off"], title_off=z["title_off"], res_type=z["res_type"],
        bfactor=z["bfactor"], xyz=z["xyz"], titles=z["titles"],
        meta=np.ascontiguousarray(z["meta"]).view(abi.META_DTYPE).reshape(-1),
    )

    class G:
        pass

    g = G()
    g.z = z
    g.batch = batch
    g.names = [str(x) for x in z["names"]]
    g.anchors = [int(x) for x in z["anchors"]]

    def blobs(b):
        off = z[f"fcz_off_{b}"]
        data = z[f"fcz_{b}"]
        return [bytes(data[int(off[i]) : int(off[i + 1])]) for i in range(len(off) - 1)]

    def decoded(b, c):
        # reference decode of chain c at anchor threshold b
        a0 = sum(_natoms(batch, i) for i in range(c))
        a1 = a0 + _natoms(batch, c)
        r0, r1 = int(batch.res_off[c]), int(batch.res_off[c + 1])
        return z[f"dec_xyz_{b}"][a0:a1], z[f"dec_bfac_{b}"][r0:r1]

    def _natoms(batch, c):
        return int(batch.atom_off[c + 1] - batch.atom_off[c])

    g.blobs = blobs
    g.decoded = decoded
    off = z["db_fcz_off"]
    g.db_blobs = [bytes(z["db_fcz"][int(off[i]) : int(off[i + 1])]) for i in range(len(off) - 1)]
    g.db_xyz = z["db_dec_xyz"]
    return g


@pytest.fixture(scope="session")
def engine():
    """The CUDA engine on cuda:0 -- fails (never falls back) when the library or the GPU is missing."""
    from foldcomp_b200.engine import Engine

    eng = Engine(0)
    yield eng
    eng.close()
