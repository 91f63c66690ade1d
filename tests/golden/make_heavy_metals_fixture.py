"""Reads the reference's Heavy_metals/processed_data.RDS (gzip + R XDR serialisation v3, written by R 3.6.3) and writes a
12 000-observation subset as tests/golden/heavy_metals_subset.npz (config 2 acceptance fixture).

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_heavy_metals_fixture.py
"""
import gzip
import os
import struct

import numpy as np

SRC = "/root/reference/Heavy_metals/processed_data.RDS"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heavy_metals_subset.npz")


class XDR:
    """minimal reader of R's serialisation format: vectors, strings, lists, pairlist attributes, REFSXP, ALTREP wrappers"""

    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def f64(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
        self.p += 8 * n
        return a

    def ints(self, n):
        a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
        self.p += 4 * n
        return a

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:                      # NILVALUE
            return None
        if t == 253 or t == 242 or t == 241:   # global env / empty env / base env
            return {"env": t}
        if t == 255:                      # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:                        # SYMSXP
            s = self.item()
            self.refs.append(s)
            return s
        if t == 9:                        # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if t == 238:                      # ALTREP: (info pairlist, state, attr)
            info, state, attr = self.item(), self.item(), self.item()
            cls = info[0][1] if isinstance(info, list) and info else None
            val = state
            if isinstance(state, list) and state and isinstance(state[0], tuple):   # pairlist state: take first value
                val = state[0][1]
            if cls == "compact_intseq" and isinstance(val, np.ndarray):
                n, start, step = int(val[0]), int(val[1]), int(val[2])
                val = np.arange(start, start + n * step, step, dtype=np.int32)
            if isinstance(val, dict) or isinstance(val, np.ndarray) or isinstance(val, list):
                return {"altrep": cls, "value": val, "attr": attr}
            return {"altrep": cls, "value": val, "attr": attr}
        if t in (2, 6):                   # LISTSXP / LANGSXP (pairlist)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nflags = self.i32()
                nt = nflags & 0xFF
                if nt == 254:
                    break
                if nt not in (2, 6):      # dotted cdr
                    self.p -= 4
                    out.append((None, self.item()))
                    break
                has_attr, has_tag = bool(nflags & 0x200), bool(nflags & 0x400)
            return out
        if t == 10:
            v = self.ints(self.length())
        elif t == 13:
            v = self.ints(self.length())
        elif t == 14:
            v = self.f64(self.length())
        elif t == 16:
            v = [self.item() for _ in range(self.length())]
        elif t == 19 or t == 20:
            v = [self.item() for _ in range(self.length())]
        else:
            raise NotImplementedError(f"SEXP type {t} at byte {self.p}")
        attr = self.item() if has_attr else None
        if attr is not None:
            return {"value": v, "attr": {k: a for k, a in attr if k is not None}}
        return v


def unwrap(x):
    while isinstance(x, dict) and "value" in x:
        x = x["value"]
    return x


def attr_of(x, name):
    while isinstance(x, dict):
        a = x.get("attr")
        if isinstance(a, list):
            a = {k: v for k, v in a if k is not None}
        if isinstance(a, dict) and name in a:
            return a[name]
        x = x.get("value")
    return None


def main():
    raw = gzip.open(SRC).read()
    assert raw[:2] == b"X\n"
    r = XDR(raw)
    r.p = 2
    version, _writer, _minver = r.i32(), r.i32(), r.i32()
    if version == 3:
        n = r.i32()
        r.p += n
    top = r.item()
    names = [unwrap(s) for s in unwrap(attr_of(top, "names"))]
    vals = unwrap(top)
    d = dict(zip(names, vals))
    locs = unwrap(d["observed_locs"]).reshape(2, -1).T            # column-major 64274 x 2 (lon, lat)
    y = unwrap(d["observed_field"])
    Xl = d["X_locs"]
    cols = unwrap(Xl)
    cnames = [unwrap(s) for s in unwrap(attr_of(Xl, "names"))]
    num, fac = {}, {}
    for nm, c in zip(cnames, cols):
        levels = attr_of(c, "levels")
        v = unwrap(c)
        if levels is not None:
            fac[nm] = (np.asarray(v, dtype=np.int32), [unwrap(s) for s in unwrap(levels)])
        else:
            num[nm] = np.asarray(v, dtype=np.float64)
    print("obs", y.size, "numeric", list(num), "factors", {k: len(v[1]) for k, v in fac.items()})
    # spatially stratified subset: every 5th observation after a deterministic shuffle keeps duplicates of locations rare but present
    rng = np.random.default_rng(0)
    sel = np.sort(rng.choice(y.size, 12000, replace=False))
    np.savez_compressed(OUT, observed_locs=locs[sel], observed_field=y[sel],
                        X_numeric=np.column_stack([num[k][sel] for k in num]), numeric_names=np.array(list(num)),
                        **{f"factor_{k}": v[0][sel] for k, v in fac.items()},
                        **{f"levels_{k}": np.array(v[1]) for k, v in fac.items()})
    print("wrote", OUT, os.path.getsize(OUT) / 1e6, "MB")


if __name__ == "__main__":
    main()
