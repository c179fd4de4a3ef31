"""CPU: the `.rds` codec and the trained-model layout (magmaclust_b200/rds_io.py).

No R interpreter exists in the image, so the codec is pinned two ways: (1) a known-answer serialisation derived by hand from
the published format (R Internals §1.8 'Serialization formats'), (2) the reference's OWN files — decoding and re-encoding
each of its eight `.rds` files must reproduce the serialised bytes exactly (runs where /root/reference exists, i.e. in the
build container; the digests of those byte streams are committed in tests/golden/rds_digests.json)."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from magmaclust_b200 import rds_io as R

HERE = os.path.dirname(__file__)
REF = "/root/reference"
DIGESTS = json.load(open(os.path.join(HERE, "golden", "rds_digests.json")))


def test_known_answer_named_double_with_na():
    """serialize(c(a = 1.5, b = NA), NULL, version = 2) as R 3.6.3 writes it"""
    want = bytes.fromhex(
        "580a" "00000002" "00030603" "00020300"
        "0000020e" "00000002" "3ff8000000000000" "7ff00000000007a2"
        "00000402" "00000001" "00040009" "00000005" + b"names".hex() +
        "00000010" "00000002" "00040009" "00000001" "61" "00040009" "00000001" "62"
        "000000fe")
    x = R.named_num([1.5, R.na_real()], ["a", "b"])
    assert R.encode(x, version=2, writer=0x00030603) == want
    back, hdr = R.decode(want)
    assert hdr == {"version": 2, "writer": 0x00030603, "min_reader": 0x00020300, "encoding": None}
    assert back.names == ["a", "b"] and back[0] == 1.5 and R.is_na_real(back)[1] and not R.is_na_real(back)[0]


def test_known_answer_symbol_reference_and_class_bit():
    """a second use of a symbol is a back reference; a class attribute sets the object bit of the header word"""
    inner = R.RList([R.to_r(1), R.to_r(True)], names=["n", "ok"])
    df = R.tibble({"x": np.array([1.0, 2.0])})
    buf = R.encode(R.RList([inner, df], names=["first", "second"]), version=2, writer=0x00030603)
    assert buf.count(b"\x00\x00\x00\x01\x00\x04\x00\x09\x00\x00\x00\x05names") == 1     # SYMSXP 'names' once ...
    assert buf.count(b"\x00\x00\x01\xff") == 2                                              # ... then REFSXP #1 twice
    assert b"\x00\x00\x03\x13\x00\x00\x00\x01" in buf                                     # VECSXP | obj | attr, length 1
    obj, _ = R.decode(buf)
    assert obj["second"].attrs["class"] == ["tbl_df", "tbl", "data.frame"]
    assert list(np.asarray(obj["second"].attrs["row.names"])) == [R.NA_INTEGER, -2]
    assert obj["first"]["ok"].rtype == R.LGLSXP and obj["first"]["n"].rtype == R.INTSXP


def test_value_model_round_trip(tmp_path):
    m = np.arange(6.0).reshape(2, 3)
    obj = {"num": 2.5, "int": np.array([1, R.NA_INTEGER, 3], dtype=np.int32), "lgl": np.array([True, False]),
           "chr": ["a", None, "é"], "mat": m, "nested": {"empty": None, "list": [1.0, "x", [2, 3]]}}
    for compress in (True, False):
        p = str(tmp_path / f"v{int(compress)}.rds")
        R.write_rds(obj, p, compress=compress)
        assert (open(p, "rb").read(2) == b"\x1f\x8b") == compress
        back = R.read_rds(p)
        assert back.names == list(obj.keys())
        assert float(back["num"][0]) == 2.5 and list(np.asarray(back["int"])) == [1, R.NA_INTEGER, 3]
        assert back["lgl"].rtype == R.LGLSXP and list(np.asarray(back["lgl"])) == [1, 0]
        assert list(back["chr"]) == ["a", None, "é"]
        assert back["mat"].dim == (2, 3) and np.array_equal(back["mat"].matrix(), m)
        assert list(np.asarray(back["mat"])) == [0.0, 3.0, 1.0, 4.0, 2.0, 5.0]                 # column-major, as R stores it
        assert back["nested"]["empty"] is None and back["nested"]["list"][1] == ["x"]
    # same bytes whichever way the file was written; a second encoding of the decoded object is identical
    assert R.encode(R.read_rds(p)) == R.encode(obj)


def test_malformed_streams_are_rejected():
    good = R.encode({"a": 1.0})
    with pytest.raises(ValueError):
        R.decode(b"A\n" + good[2:])
    with pytest.raises(ValueError):
        R.decode(good[:-3])
    with pytest.raises(ValueError):
        R.decode(good + b"\x00")
    with pytest.raises(TypeError):
        R.encode({"a": object()})


def _fake_trained(K=3, M=4, N=5, seed=0):
    rng = np.random.default_rng(seed)
    t = np.array([0.0, 0.1, 0.25, 1.0 / 3.0, 10.0])[:N]
    names_k = [f"K{k + 1}" for k in range(K)]
    ids = [str(i + 2) for i in range(M)]

    class Named:                      # stands in for magma.NamedMat (.m, .t)
        def __init__(self, m, t):
            self.m, self.t = m, t
    cov = {}
    for k in names_k:
        a = rng.standard_normal((N, N))
        cov[k] = Named(a @ a.T, t)
    tau = rng.dirichlet(np.ones(K), M)
    return {"hp": {"theta_k": {k: rng.standard_normal(3) for k in names_k},
                   "theta_i": {i: rng.standard_normal(3) for i in ids}, "pi_k": tau.mean(0)},
            "convergence": True,
            "param": {"mean": {k: {"Timestamp": t, "Output": rng.standard_normal(N)} for k in names_k}, "cov": cov,
                      "tau_i_k": {k: {i: float(tau[r, c]) for r, i in enumerate(ids)} for c, k in enumerate(names_k)}},
            "Time_train": 1.25, "logL_iterations": np.array([-10.0, -9.0, -8.99])}


def test_trained_model_layout_and_round_trip(tmp_path):
    """MC.R:94-98: list(hp, convergence, param, Time_train, logL_iterations) with the attributes R code relies on"""
    res = _fake_trained()
    p = str(tmp_path / "trained.rds")
    R.save_trained(res, p)
    obj = R.read_rds(p)
    assert obj.names == ["hp", "convergence", "param", "Time_train", "logL_iterations"]
    assert obj["hp"].names == ["theta_k", "theta_i", "pi_k"] and obj["hp"]["pi_k"].names == ["K1", "K2", "K3"]
    assert obj["hp"]["theta_k"]["K2"].names == ["p1", "p2", ""]            # CF.R:666-668 appends the unnamed pen_diag
    assert obj["Time_train"].attrs["class"] == ["difftime"] and obj["Time_train"].attrs["units"] == ["secs"]
    mean = obj["param"]["mean"]["K1"]
    assert mean.names == ["Timestamp", "Output"] and mean.attrs["class"] == ["tbl_df", "tbl", "data.frame"]
    cov = obj["param"]["cov"]["K3"]
    assert cov.dim == (5, 5)
    assert list(cov.attrs["dimnames"][0]) == ["X0", "X0.1", "X0.25", "X0.333333333333333", "X10"]   # paste0('X', t), CF.R:83
    assert list(cov.attrs["dimnames"][1]) == list(cov.attrs["dimnames"][0])
    assert obj["param"]["tau_i_k"]["K1"].names == ["2", "3", "4", "5"]
    back = R.load_trained(p)
    for k in res["hp"]["theta_k"]:
        assert np.array_equal(back["hp"]["theta_k"][k], res["hp"]["theta_k"][k])
        assert np.array_equal(back["param"]["mean"][k]["Output"], res["param"]["mean"][k]["Output"])
        assert np.array_equal(back["param"]["cov"][k], res["param"]["cov"][k].m)
        assert back["param"]["tau_i_k"][k] == res["param"]["tau_i_k"][k]
    for i in res["hp"]["theta_i"]:
        assert np.array_equal(back["hp"]["theta_i"][i], res["hp"]["theta_i"][i])
    assert np.array_equal(back["hp"]["pi_k"], res["hp"]["pi_k"]) and back["convergence"] is True
    assert back["Time_train"] == 1.25 and np.array_equal(back["logL_iterations"], res["logL_iterations"])


needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")


@needs_ref
@pytest.mark.parametrize("rel", sorted(DIGESTS))
def test_reference_files_reencode_byte_exact(rel):
    """decode -> encode of each `.rds` file the reference ships reproduces R's own serialisation byte for byte"""
    raw = gzip.decompress(open(os.path.join(REF, rel), "rb").read())
    assert hashlib.sha256(raw).hexdigest() == DIGESTS[rel]["sha256"] and len(raw) == DIGESTS[rel]["bytes"]
    obj, hdr = R.decode(raw)
    assert hdr == DIGESTS[rel]["header"]
    again = R.encode(obj, version=hdr["version"], writer=hdr["writer"], min_reader=hdr["min_reader"],
                     encoding=hdr["encoding"] or "UTF-8")
    assert hashlib.sha256(again).hexdigest() == DIGESTS[rel]["sha256"]


@needs_ref
def test_study_layout_loads_like_the_golden_fixture():
    """`[[d]]$MAGMAclust` of the stored study outputs (hp with `opm` data frames, tau_i_k at top level) through
    `trained_from_r` == the arrays tests/golden/make_golden.py committed from the same file"""
    gold = np.load(os.path.join(HERE, "golden", "hoo_diff_time.npz"))
    rds = R.read_rds(os.path.join(REF, "Simulations/Training/train_for_pred_Hoo_diff_time.rds"))
    for d in (0, 57, 99):
        got = R.trained_from_r(rds[d]["MAGMAclust"])
        assert list(got["hp"]["theta_k"]) == ["K1", "K2", "K3"] and list(got["hp"]["theta_i"]) == [str(i) for i in range(2, 22)]
        assert np.array_equal(np.array(list(got["hp"]["theta_k"].values())), gold["theta_k"][d])
        assert np.array_equal(np.array(list(got["hp"]["theta_i"].values())), gold["theta_i"][d])
        assert np.array_equal(got["hp"]["pi_k"], gold["pi_k"][d])
        tau = np.array([[got["param"]["tau_i_k"][k][i] for k in ("K1", "K2", "K3")] for i in got["hp"]["theta_i"]])
        assert np.array_equal(tau, gold["tau_i_k"][d])
        assert got["Time_train"] == gold["Time_train"][d]


def test_magma_save_and_load_trained(tmp_path):
    """the host mirror's entry points: IDs converted back to the type of db['ID'], covariances as NamedMat"""
    from magmaclust_b200 import magma
    res = _fake_trained()
    p = str(tmp_path / "m.rds")
    magma.save_trained(res, p)
    back = magma.load_trained(p, id_type=int)
    assert list(back["hp"]["theta_i"]) == [2, 3, 4, 5] and list(back["param"]["tau_i_k"]["K2"]) == [2, 3, 4, 5]
    c = back["param"]["cov"]["K1"]
    assert isinstance(c, magma.NamedMat) and np.array_equal(c.m, res["param"]["cov"]["K1"].m)
    assert np.array_equal(c.sub([0.1, 10.0]), res["param"]["cov"]["K1"].m[np.ix_([1, 4], [1, 4])])
    # an already decoded list is accepted as well (the study keeps one list per dataset in a single file)
    again = magma.load_trained(R.read_rds(p), id_type=int)
    assert np.array_equal(again["hp"]["pi_k"], back["hp"]["pi_k"])


def test_random_objects_round_trip():
    """property test: any nesting of the supported value types survives encode -> decode -> encode unchanged"""
    from hypothesis import given, settings, strategies as st

    finite = st.floats(allow_nan=False, allow_infinity=True, width=64)
    leaf = st.one_of(
        st.none(),
        st.lists(finite, max_size=5).map(lambda v: R.RVector(np.array(v, dtype=np.float64), R.REALSXP)),
        st.lists(st.integers(-2 ** 31, 2 ** 31 - 1), max_size=5).map(lambda v: R.RVector(np.array(v, dtype=np.int64), R.INTSXP)),
        st.lists(st.booleans(), max_size=4).map(lambda v: R.RVector(np.array(v, dtype=bool), R.LGLSXP)),
        st.lists(st.one_of(st.none(), st.text(max_size=6)), max_size=4).map(R.RStrings))
    names = st.text(alphabet="abcXYZ_.12", min_size=1, max_size=5)

    def named(children):
        return st.lists(st.tuples(names, children), max_size=4).map(
            lambda kv: R.RList([v for _, v in kv], names=[k for k, _ in kv]) if kv else R.RList([]))

    tree = st.recursive(leaf, lambda ch: st.one_of(named(ch), st.lists(ch, max_size=3).map(R.RList)), max_leaves=12)

    @settings(max_examples=150, deadline=None)
    @given(tree, st.sampled_from([2, 3]))
    def check(obj, version):
        buf = R.encode(obj, version=version)
        back, hdr = R.decode(buf)
        assert hdr["version"] == version
        assert R.encode(back, version=version) == buf

    check()
