"""The native GMNS reader (dtl_gmns_read) against the reference's own in-memory network image
(g_read_input_data + connectors, dumped by oracle/_ref/dtalite_ref; digests in tests/golden).  CPU only."""
import hashlib
import os

import numpy as np
import pytest

from conftest import FIXTURES


@pytest.mark.parametrize("name", ["two_corridor", "corridor_101", "corridor_3", "corridor_3_2p", "edge_cases", "toll_signal", "asu_qvdf"])
def test_reader_image_is_bit_identical_to_reference(problems, golden, name):
    p, g = problems(name), golden(name)
    N, L, Z, AT, T = (int(x) for x in g["meta"])
    assert (p.n_nodes, p.n_links, p.n_zones, p.n_agent_types, p.n_periods) == (N, L, Z, AT, T)
    assert np.float32(p.total_demand_volume) == np.float32(g["total_demand_volume"])
    for name_, digest in zip(g["problem_array_names"], g["problem_array_sha256"]):
        got = hashlib.sha256(np.ascontiguousarray(p.arrays[str(name_)]).tobytes()).hexdigest()
        assert got == str(digest), f"array {name_} differs from the reference's image"


def test_reader_structure_two_corridor(problems):
    p = problems("two_corridor")
    # 4 physical nodes then centroids of zones 1, 2 (node_id = -zone_id); connectors appended per activity node
    assert list(p.node_id) == [1, 2, 3, 4, -1, -2]
    assert list(p.zone_id) == [1, 2]
    assert list(p.link_tag) == [-1, -1, -1, -1, -1, 0, -1, 1]
    assert list(p.link_from[4:]) == [0, 4, 1, 5] and list(p.link_to[4:]) == [4, 0, 5, 1]
    assert np.all(p.fftt[0, :4] == [10, 10, 15, 15]) and np.all(p.fftt[0, 4:] == 1e-4)
    assert np.all(p.qdf[0, 4:] == -1) and np.all(p.vdf_type[0, :4] == 1) and np.all(p.vdf_type[0, 4:] == 0)


def test_reader_rejects_unsupported_inputs(native, tmp_path):
    import shutil
    import dlsim_mrm_b200 as D
    src = os.path.join(FIXTURES, "two_corridor")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp_path)
    s = open(tmp_path / "settings.csv").read().replace(",1,demand.csv,,column,AM,p,", ",1,demand.csv,,matrix,AM,p,")
    open(tmp_path / "settings.csv", "w").write(s)
    with pytest.raises(D.EngineError, match="not supported"):
        D.read_gmns(str(tmp_path))
    os.remove(tmp_path / "node.csv")
    with pytest.raises(D.EngineError, match="node.csv"):
        D.read_gmns(str(tmp_path))


def test_assignment_settings(native):
    import ctypes as C
    out = (C.c_int * 6)()
    assert native.dtl_gmns_assignment_settings(os.path.join(FIXTURES, "two_corridor").encode(), out) == 0
    assert list(out) == [20, 0, 0, 0, 0, 4]


def test_parallel_link_parse_gives_the_same_image(native, monkeypatch):
    """link.csv is cut into slices of whole lines that worker threads parse; the image must not depend on
    the number of slices -- including a count that does not divide the file evenly"""
    import dlsim_mrm_b200 as D
    d = os.path.join(FIXTURES, "corridor_101")          # link.csv: 2 MB, above the 1 MiB threshold of the parallel path
    assert os.path.getsize(os.path.join(d, "link.csv")) > (1 << 20)
    images = []
    for threads in ("1", "7", "16"):
        monkeypatch.setenv("DTL_IO_THREADS", threads)
        images.append(D.read_gmns(d, 4))
    for other in images[1:]:
        assert other.n_links == images[0].n_links and other.n_nodes == images[0].n_nodes
        for name, a in images[0].arrays.items():
            assert np.array_equal(a, other.arrays[name], equal_nan=a.dtype.kind == "f"), name
