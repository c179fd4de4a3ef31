"""R `.rds` files (gzip + XDR serialisation, format version 2 or 3): reader, writer and the trained-model layout.

SURVEY.md §8f row 4 (the data formats either side of the path): the study scripts keep their trained models as
`saveRDS(list(...))` (Simulation_study.R:554-582, `Simulations/Training/*.rds`) and read them back with `readRDS`. This
module lets the GPU path exchange those files without R:

* `read_rds` / `write_rds` — a LOSSLESS codec for the SEXP types that occur in such files (NULL, logical, integer,
  double, character, generic vectors, attribute pairlists, symbols and symbol references, ALTREP compact sequences —
  expanded on reading, written back compact while the values still are the sequence). Lossless means: decoding one of
  the reference's own files and encoding the result reproduces its serialised bytes exactly; tests/test_rds_io.py does
  that for all eight `.rds` files of the reference (`Simulations/Training/*.rds`, `Real_Data_Study/Training/*.rds`),
  which is the strongest check available without an R interpreter in the image.
* `trained_to_r` / `trained_from_r` — `training_VEM`'s return value (MC.R:94-98: hp, convergence, param, Time_train,
  logL_iterations) between the Python dict layout of `magma.py` and the R list layout, including the layout the study
  stores (`[[d]]$MAGMAclust`: hp with `opm` data frames for theta_i, Time_train, tau_i_k at top level).

Host-side format code only: nothing numerical happens here.
"""
from __future__ import annotations

import gzip
import struct
from collections import OrderedDict

import numpy as np

NILVALUE, REFSXP, ALTREP = 254, 255, 238
SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 1, 2, 9, 10, 13, 14, 16, 19
NA_INTEGER = -2147483648
_NA_REAL_BITS = 0x7FF00000000007A2
F_OBJ, F_ATTR, F_TAG = 1 << 8, 1 << 9, 1 << 10
GP_LATIN1, GP_UTF8, GP_ASCII = 1 << 2, 1 << 3, 1 << 6


def na_real() -> float:
    """R's NA_real_ (a NaN with payload 1954)"""
    return float(np.array([_NA_REAL_BITS], dtype=np.uint64).view(np.float64)[0])


def is_na_real(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.float64).view(np.uint64) == _NA_REAL_BITS


class _WithAttrs:
    """R attributes in file order (`names` included); `gp` = the general-purpose bits of the header word"""

    attrs = None
    gp = 0

    @property
    def names(self):
        n = (self.attrs or {}).get("names")
        return None if n is None else list(n)

    @names.setter
    def names(self, value):
        if self.attrs is None:
            self.attrs = OrderedDict()
        if value is None:
            self.attrs.pop("names", None)
        else:
            self.attrs["names"] = value if isinstance(value, RStrings) else RStrings(value)


class RList(_WithAttrs, list):
    """generic vector (VECSXP); indexable by position or by name"""

    def __init__(self, items=(), names=None, attrs=None):
        list.__init__(self, items)
        self.attrs = OrderedDict(attrs or {})
        if names is not None:
            self.names = names

    def __getitem__(self, key):
        if isinstance(key, str):
            return list.__getitem__(self, self.names.index(key))
        return list.__getitem__(self, key)

    def keys(self):
        return self.names or []

    def as_dict(self):
        return OrderedDict(zip(self.names, self))


class RStrings(_WithAttrs, list):
    """character vector (STRSXP): list of str, None = NA_character_"""

    def __init__(self, items=(), attrs=None):
        list.__init__(self, [items] if isinstance(items, str) else items)
        self.attrs = OrderedDict(attrs or {})


class RVector(_WithAttrs, np.ndarray):
    """logical / integer / double vector: numpy array (int32 for logical and integer, NA = NA_INTEGER) + `rtype`"""

    rtype = REALSXP
    altrep = None

    def __new__(cls, arr, rtype=None, attrs=None):
        a = np.asarray(arr)
        if rtype is None:
            rtype = LGLSXP if a.dtype == np.bool_ else INTSXP if np.issubdtype(a.dtype, np.integer) else REALSXP
        a = a.astype(np.float64 if rtype == REALSXP else np.int32, copy=False)
        obj = np.ascontiguousarray(a).view(cls)
        obj.rtype = rtype
        obj.attrs = OrderedDict(attrs or {})
        return obj

    def __array_finalize__(self, obj):
        self.rtype = getattr(obj, "rtype", REALSXP)
        self.attrs = OrderedDict(getattr(obj, "attrs", None) or {})
        self.gp = getattr(obj, "gp", 0)

    @property
    def dim(self):
        d = (self.attrs or {}).get("dim")
        return None if d is None else tuple(int(x) for x in np.asarray(d))

    def matrix(self) -> np.ndarray:
        """plain C-ordered numpy array of an R matrix / array (R stores column-major)"""
        d = self.dim
        return np.asarray(self) if d is None else np.asarray(self).reshape(d[::-1]).transpose()


def _expand_compact(cls, state):
    """the vector an ALTREP compact sequence stands for (state = c(length, first, step) as doubles)"""
    if cls == "compact_intseq":
        n, start, step = (int(x) for x in np.asarray(state)[:3])
        return RVector(start + step * np.arange(n, dtype=np.int64), INTSXP)
    if cls == "compact_realseq":
        n, start, step = (float(x) for x in np.asarray(state)[:3])
        return RVector(start + step * np.arange(int(n)), REALSXP)
    return None


# ---- reader ---------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.o, self.refs = buf, 0, []

    def i32(self) -> int:
        if self.o + 4 > len(self.b):
            raise ValueError("truncated rds stream")
        (v,) = struct.unpack_from(">i", self.b, self.o)
        self.o += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        if len(v) != n:
            raise ValueError("truncated rds stream")
        self.o += n
        return v

    def attributes(self) -> OrderedDict:
        out = OrderedDict()
        pl = self.item()
        for tag, val in pl or []:
            out[tag] = val
        return out

    def pairlist(self, flags):
        out = []
        while True:
            if flags & F_ATTR:
                self.item()
            tag = self.item() if flags & F_TAG else None
            out.append((tag, self.item()))
            flags = self.i32()
            ty = flags & 0xFF
            if ty == NILVALUE:
                return out
            if ty != LISTSXP:
                raise ValueError(f"unexpected pairlist cdr type {ty}")

    def item(self):
        flags = self.i32()
        ty = flags & 0xFF
        if ty == NILVALUE:
            return None
        if ty == REFSXP:
            idx = flags >> 8
            return self.refs[(idx if idx else self.i32()) - 1]
        if ty == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if ty == LISTSXP:
            return self.pairlist(flags)
        if ty == CHARSXP:
            n = self.i32()
            if n == -1:
                return None
            return self.raw(n).decode("latin-1" if (flags >> 12) & GP_LATIN1 else "utf-8", "replace")
        if ty == ALTREP:
            info, state = self.item(), self.item()
            attrs = self.attributes()
            cls = info[0][1] if isinstance(info, list) else None
            val = _expand_compact(cls, state)
            if val is None:
                raise ValueError(f"unsupported ALTREP class {cls}")
            val.attrs = attrs
            val.altrep = (cls, info[1][1], int(np.asarray(info[2][1])[0]), state)   # class, package, base type, state
            return val
        if ty in (LGLSXP, INTSXP):
            n = self.i32()
            val = RVector(np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32), ty)
        elif ty == REALSXP:
            n = self.i32()
            val = RVector(np.frombuffer(self.raw(8 * n), dtype=">u8").astype(np.uint64).view(np.float64), REALSXP)
        elif ty == STRSXP:
            n = self.i32()
            val = RStrings(self.item() for _ in range(n))
        elif ty == VECSXP:
            n = self.i32()
            val = RList(self.item() for _ in range(n))
        else:
            raise ValueError(f"unsupported SEXP type {ty} at offset {self.o - 4}")
        val.gp = (flags >> 12) & 0xFFFF
        if flags & F_ATTR:
            val.attrs = self.attributes()
        return val


def decode(buf: bytes):
    """(object, header) of an uncompressed XDR serialisation; header = dict(version, writer, min_reader, encoding)"""
    if buf[:2] != b"X\n":
        raise ValueError("not an XDR-serialised R object")
    r = _Reader(buf)
    r.o = 2
    hdr = {"version": r.i32(), "writer": r.i32(), "min_reader": r.i32(), "encoding": None}
    if hdr["version"] not in (2, 3):
        raise ValueError(f"unsupported serialisation version {hdr['version']}")
    if hdr["version"] == 3:
        hdr["encoding"] = r.raw(r.i32()).decode("ascii")
    obj = r.item()
    if r.o != len(buf):
        raise ValueError("trailing bytes after the serialised object")
    return obj, hdr


def read_rds(path: str, with_header: bool = False):
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:2] == b"\x1f\x8b":
        buf = gzip.decompress(buf)
    obj, hdr = decode(buf)
    return (obj, hdr) if with_header else obj


# ---- writer ---------------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.out, self.syms = [], {}

    def i32(self, v: int):
        self.out.append(struct.pack(">i", v))

    def charsxp(self, s):
        if s is None:
            self.i32(CHARSXP)
            self.i32(-1)
            return
        try:
            b, gp = s.encode("ascii"), GP_ASCII
        except UnicodeEncodeError:
            b, gp = s.encode("utf-8"), GP_UTF8
        self.i32(CHARSXP | (gp << 12))
        self.i32(len(b))
        self.out.append(b)

    def symbol(self, name: str):
        idx = self.syms.get(name)
        if idx is not None:
            if idx < (1 << 23):
                self.i32((idx << 8) | REFSXP)
            else:
                self.i32(REFSXP)
                self.i32(idx)
            return
        self.syms[name] = len(self.syms) + 1
        self.i32(SYMSXP)
        self.charsxp(name)

    def attributes(self, attrs):
        for tag, val in attrs.items():
            self.i32(LISTSXP | F_TAG)
            self.symbol(tag)
            self.item(val)
        self.i32(NILVALUE)

    def item(self, x):
        if x is None:
            self.i32(NILVALUE)
            return
        x = to_r(x)
        attrs = x.attrs or {}
        flags = (x.gp << 12) | (F_ATTR if attrs else 0) | (F_OBJ if "class" in attrs else 0)
        alt = getattr(x, "altrep", None) if isinstance(x, RVector) else None
        if alt is not None:
            # a compact sequence read from a file goes back as one, as long as the values still are that sequence
            ref = _expand_compact(alt[0], alt[3])
            if ref is not None and ref.rtype == x.rtype and ref.shape == x.shape and np.array_equal(np.asarray(ref), np.asarray(x)):
                self.i32(ALTREP | flags)
                self.i32(LISTSXP)
                self.symbol(alt[0])
                self.i32(LISTSXP)
                self.symbol(alt[1])
                self.i32(LISTSXP)
                self.item(RVector(np.array([alt[2]]), INTSXP))
                self.i32(NILVALUE)
                self.item(alt[3])
                if attrs:
                    self.attributes(attrs)
                else:
                    self.i32(NILVALUE)
                return
        if isinstance(x, RVector):
            self.i32(x.rtype | flags)
            self.i32(x.size)
            a = np.asarray(x)
            if x.rtype == REALSXP:
                self.out.append(np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).astype(">u8").tobytes())
            else:
                self.out.append(np.ascontiguousarray(a, dtype=np.int32).astype(">i4").tobytes())
        elif isinstance(x, RStrings):
            self.i32(STRSXP | flags)
            self.i32(len(x))
            for s in x:
                self.charsxp(s)
        elif isinstance(x, RList):
            self.i32(VECSXP | flags)
            self.i32(len(x))
            for v in x:
                self.item(v)
        else:
            raise TypeError(f"cannot serialise {type(x).__name__}")
        if attrs:
            self.attributes(attrs)


def to_r(x):
    """Python value -> R value model: scalars become length-1 vectors, str / list of str a character vector, dict a
    named list, list / tuple an unnamed list, 2-D arrays a matrix (column-major data + dim)"""
    if x is None or isinstance(x, (RVector, RStrings, RList)):
        return x
    if isinstance(x, str):
        return RStrings([x])
    if isinstance(x, (bool, np.bool_)):
        return RVector(np.array([x]), LGLSXP)
    if isinstance(x, (int, np.integer)):
        return RVector(np.array([x]), INTSXP)
    if isinstance(x, (float, np.floating)):
        return RVector(np.array([x]), REALSXP)
    if isinstance(x, dict):
        return RList([to_r(v) for v in x.values()], names=[str(k) for k in x.keys()])
    if isinstance(x, (list, tuple)):
        if len(x) and all(isinstance(s, str) or s is None for s in x):
            return RStrings(x)
        return RList([to_r(v) for v in x])
    if isinstance(x, np.ndarray):
        if x.dtype.kind in "US":
            return RStrings([str(s) for s in x.ravel()])
        if x.ndim <= 1:
            return RVector(x.ravel())
        v = RVector(np.ascontiguousarray(np.transpose(x)).ravel())
        v.attrs["dim"] = RVector(np.array(x.shape), INTSXP)
        return v
    raise TypeError(f"no R equivalent for {type(x).__name__}")


def encode(obj, version: int = 3, writer: int = 0x00040100, min_reader: int | None = None, encoding: str = "UTF-8") -> bytes:
    w = _Writer()
    w.out.append(b"X\n")
    w.i32(version)
    w.i32(writer)
    w.i32(min_reader if min_reader is not None else (0x00030500 if version == 3 else 0x00020300))
    if version == 3:
        e = encoding.encode("ascii")
        w.i32(len(e))
        w.out.append(e)
    elif version != 2:
        raise ValueError("serialisation version must be 2 or 3")
    w.item(obj)
    return b"".join(w.out)


def write_rds(obj, path: str, compress: bool = True, **header):
    """`saveRDS(obj, path)`: gzip-compressed by default like R's; `header` = version / writer / min_reader / encoding"""
    buf = encode(obj, **header)
    with open(path, "wb") as f:
        if compress:
            with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:
                g.write(buf)
        else:
            f.write(buf)


# ---- R idioms ---------------------------------------------------------------------------------------------------------
def named_num(values, names) -> RVector:
    v = RVector(np.asarray(values, dtype=np.float64), REALSXP)
    v.attrs["names"] = RStrings(names)
    return v


def tibble(columns: dict) -> RList:
    """a tibble as R serialises it: names, compact row.names c(NA, -n), class c('tbl_df', 'tbl', 'data.frame')"""
    cols = [to_r(v) for v in columns.values()]
    n = len(cols[0]) if cols else 0
    out = RList(cols)
    out.attrs["class"] = RStrings(["tbl_df", "tbl", "data.frame"])
    out.attrs["row.names"] = RVector(np.array([NA_INTEGER, -n]), INTSXP)
    out.attrs["names"] = RStrings([str(k) for k in columns.keys()])
    return out


def r_num_label(t: float) -> str:
    """as.character(t) for a double: 15 significant digits (the 'X<t>' dimnames of CF.R:83,146)"""
    return format(float(t), ".15g")


def named_matrix(m, t) -> RVector:
    """symmetric matrix with dimnames list(paste0('X', t), paste0('X', t)) as kern_to_cov / kern_to_inv name theirs"""
    v = to_r(np.asarray(m, dtype=np.float64))
    labels = ["X" + r_num_label(x) for x in np.asarray(t, dtype=np.float64)]
    v.attrs["dimnames"] = RList([RStrings(labels), RStrings(labels)])
    return v


# ---- training_VEM's return value (MC.R:94-98) -----------------------------------------------------------------------------
def trained_to_r(res: dict) -> RList:
    """the dict `magma.training_VEM` returns -> the R list `training_VEM` returns: hp (theta_k: list of named num(3) with
    names c('p1', 'p2', '') as CF.R:666-668 leaves them, theta_i: list of named num(3), pi_k: named num), convergence,
    param (mean: list of tibbles Timestamp / Output, cov: list of named matrices, tau_i_k: list over clusters of lists
    over individuals), Time_train (difftime in seconds), logL_iterations"""
    hp, param = res["hp"], res["param"]
    names_k = [str(k) for k in hp["theta_k"].keys()]
    r_hp = RList([
        RList([named_num(np.asarray(v)[:3], ["p1", "p2", ""]) for v in hp["theta_k"].values()], names=names_k),
        RList([named_num(np.asarray(v)[:3], ["p1", "p2", "p3"]) for v in hp["theta_i"].values()],
              names=[str(i) for i in hp["theta_i"].keys()]),
        named_num(np.asarray(hp["pi_k"]), names_k)], names=["theta_k", "theta_i", "pi_k"])
    items = [RList([tibble({"Timestamp": np.asarray(m["Timestamp"], dtype=np.float64),
                            "Output": np.asarray(m["Output"], dtype=np.float64)}) for m in param["mean"].values()],
                   names=names_k)]
    pnames = ["mean"]
    if "cov" in param:
        items.append(RList([named_matrix(getattr(c, "m", c), getattr(c, "t", next(iter(param["mean"].values()))["Timestamp"]))
                            for c in param["cov"].values()], names=names_k))
        pnames.append("cov")
    items.append(RList([RList([to_r(float(x)) for x in per_i.values()], names=[str(i) for i in per_i.keys()])
                        for per_i in param["tau_i_k"].values()], names=names_k))
    pnames.append("tau_i_k")
    tt = RVector(np.array([float(res.get("Time_train", 0.0))]), REALSXP)
    tt.attrs["class"] = RStrings(["difftime"])
    tt.attrs["units"] = RStrings(["secs"])
    out = RList([r_hp, to_r(bool(res.get("convergence", False))), RList(items, names=pnames), tt],
                names=["hp", "convergence", "param", "Time_train"])
    if "logL_iterations" in res:
        out.append(RVector(np.asarray(res["logL_iterations"], dtype=np.float64), REALSXP))
        out.names = out.names + ["logL_iterations"]
    return out


def _theta_row(v) -> np.ndarray:
    """named num(3) or the one-row `opm` data frame optimr::opm leaves in the stored study outputs (p1, p2[, p3] columns)"""
    if isinstance(v, RList):
        return np.array([float(np.asarray(col)[0]) for col in v], dtype=np.float64)
    return np.asarray(v, dtype=np.float64).copy()


def trained_from_r(obj: RList) -> dict:
    """inverse of `trained_to_r`; also accepts the layout the study stores per dataset (`[[d]]$MAGMAclust`, SS.R:554-582:
    hp, Time_train and tau_i_k at top level, theta_i as `opm` data frames). Cluster and individual names are kept as
    strings, exactly as R has them"""
    names = obj.names or []
    hp = obj["hp"]
    names_k = hp["theta_k"].names
    out_hp = {"theta_k": OrderedDict((k, _theta_row(hp["theta_k"][k])) for k in names_k),
              "theta_i": OrderedDict((i, _theta_row(v)) for i, v in zip(hp["theta_i"].names, hp["theta_i"])),
              "pi_k": np.asarray(hp["pi_k"], dtype=np.float64).copy()}
    out = {"hp": out_hp}
    param = obj["param"] if "param" in names else None
    tau = param["tau_i_k"] if param is not None and "tau_i_k" in (param.names or []) else obj["tau_i_k"] if "tau_i_k" in names else None
    p = {}
    if param is not None and "mean" in (param.names or []):
        p["mean"] = OrderedDict((k, {"Timestamp": np.asarray(m["Timestamp"], dtype=np.float64).copy(),
                                     "Output": np.asarray(m["Output"], dtype=np.float64).copy()})
                                for k, m in zip(param["mean"].names, param["mean"]))
    if param is not None and "cov" in (param.names or []):
        p["cov"] = OrderedDict((k, c.matrix().copy()) for k, c in zip(param["cov"].names, param["cov"]))
    if tau is not None:
        p["tau_i_k"] = OrderedDict((k, OrderedDict((i, float(np.asarray(x)[0])) for i, x in zip(per_i.names, per_i)))
                                   for k, per_i in zip(tau.names, tau))
    if p:
        out["param"] = p
    if "convergence" in names:
        out["convergence"] = bool(np.asarray(obj["convergence"])[0])
    if "Time_train" in names:
        out["Time_train"] = float(np.asarray(obj["Time_train"])[0])
    if "logL_iterations" in names:
        out["logL_iterations"] = np.asarray(obj["logL_iterations"], dtype=np.float64).copy()
    return out


def save_trained(res: dict, path: str, **header):
    """saveRDS(training_VEM(...), path)"""
    write_rds(trained_to_r(res), path, **header)


def load_trained(path: str) -> dict:
    """readRDS(path) of a `training_VEM` result (either layout `trained_from_r` accepts)"""
    return trained_from_r(read_rds(path))
