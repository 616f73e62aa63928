"""Reader/writer for RSGV golden-vector containers (named uint64 arrays). C++ twin: oracle/rsgv_io.hpp."""
import numpy as np


def load(path):
    raw = np.fromfile(path, dtype=np.uint8)
    assert raw[:8].tobytes() == b"RSGV0001", "bad magic"
    words = raw[8:].view(np.uint64)
    n, pos, out = int(words[0]), 1, {}
    for _ in range(n):
        ln = int(words[pos]); pos += 1
        nw_name = (ln + 7) // 8
        name = words[pos:pos + nw_name].tobytes()[:ln].decode(); pos += nw_name
        cnt = int(words[pos]); pos += 1
        out[name] = words[pos:pos + cnt].copy(); pos += cnt
    return out


def save(path, entries):
    chunks = [b"RSGV0001", np.uint64(len(entries)).tobytes()]
    for name, arr in entries.items():
        b = name.encode()
        chunks.append(np.uint64(len(b)).tobytes())
        chunks.append(b + b"\0" * ((-len(b)) % 8))
        arr = np.ascontiguousarray(arr, dtype=np.uint64)
        chunks.append(np.uint64(arr.size).tobytes())
        chunks.append(arr.tobytes())
    with open(path, "wb") as f:
        f.write(b"".join(chunks))


class Case:
    """Typed view over one dump produced by `oracle/_ref/ref_harness dump` (layout: oracle/ref_harness.cpp)."""

    def __init__(self, path):
        self.d = d = load(path)
        (self.N_R, self.L_R, self.N_E, self.L_E, self.n, self.io, self.aux, self.seed,
         self.use_const, self.quirks) = (int(x) for x in d["params"])
        self.q = d["ring_q"].copy()
        self.Q = d["enc_Q"].copy()
        self.W = self.N_R * self.L_R
        self.enc_words = self.L_R * 2 * self.L_E * self.N_E

    def ring(self, name):
        """-> (words [k][L_R*N_R], tag [k] (0 scalar / 1 poly), scalar [k])"""
        w = self.d[name].reshape(-1, self.W)
        return w, self.d[name + ".tag"], self.d[name + ".scalar"]

    def enc(self, name):
        """-> (words [k][L_R*2*L_E*N_E], size [k][L_R])"""
        return self.d[name].reshape(-1, self.enc_words), self.d[name + ".size"].reshape(-1, self.L_R)
