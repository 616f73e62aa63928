"""The encoding container (csrc/serialize.inl): host-only round trips and every rejection the reader promises (CPU tier),
and HBM <-> file round trips of a CRS range and of a proof (GPU tier).  The reference has no format to compare with
(r1cs_ppzksnark.hpp:43-47 declares operator<< / operator>> and never defines them): parity here is round-trip identity."""
import os
import sys

import numpy as np
import pytest

from ringsnark_b200 import serialize as S
from ringsnark_b200.capi import RsgError

N_R, N_E = 64, 256
q = [786433]
Q = [12289 * 0 + 1099511627777 * 0 + 576460752303439873, 576460752303702017]   # any values: the file layer only checks ranges


def _words(n, seed=1):
    rng = np.random.default_rng(seed)
    per_row = N_E
    rows = []
    for _ in range(n * len(q) * 2):
        for p in Q:
            rows.append(rng.integers(0, p, size=per_row, dtype=np.uint64))
    return np.concatenate(rows).reshape(n, -1)


def test_round_trip_and_header(tmp_path):
    w = _words(3)
    path = tmp_path / "proof.rsgk"
    S.write_encodings(path, w, N_R, q, N_E, Q, kind=S.FILE_PROOF)
    info = S.file_info(path)
    assert info == dict(kind=2, N_R=N_R, L_R=1, N_E=N_E, L_E=2, n_elems=3, q=q, Q=Q)
    got, kind = S.read_encodings(path, N_R, q, N_E, Q)
    assert kind == S.FILE_PROOF and np.array_equal(got, w)
    assert os.path.getsize(path) == 8 + 6 * 8 + 3 * 8 + w.size * 8 + 16


def test_empty_file_of_zero_encodings(tmp_path):
    path = tmp_path / "empty.rsgk"
    S.write_encodings(path, np.zeros((0, 2 * 2 * N_E), dtype=np.uint64), N_R, q, N_E, Q, kind=S.FILE_CRS)
    got, kind = S.read_encodings(path, N_R, q, N_E, Q)
    assert kind == S.FILE_CRS and got.shape[0] == 0


def test_reader_rejections(tmp_path):
    w = _words(2)
    path = tmp_path / "crs.rsgk"
    S.write_encodings(path, w, N_R, q, N_E, Q, kind=S.FILE_CRS)
    raw = bytearray(open(path, "rb").read())
    # other parameters
    with pytest.raises(RsgError, match="other parameters"):
        S.read_encodings(path, N_R, q, N_E, [Q[0], Q[1] + 2])
    with pytest.raises(RsgError, match="other parameters"):
        S.read_encodings(path, N_R * 2, q, N_E, Q)
    # flipped payload bit -> checksum
    bad = bytearray(raw)
    bad[8 + 9 * 8 + 5] ^= 1
    (tmp_path / "flip.rsgk").write_bytes(bad)
    with pytest.raises(RsgError, match="checksum"):
        S.read_encodings(tmp_path / "flip.rsgk", N_R, q, N_E, Q)
    # truncated
    (tmp_path / "short.rsgk").write_bytes(raw[:len(raw) // 2])
    with pytest.raises(RsgError, match="short file"):
        S.read_encodings(tmp_path / "short.rsgk", N_R, q, N_E, Q)
    # wrong magic
    bad = bytearray(raw)
    bad[0] = ord("X")
    (tmp_path / "magic.rsgk").write_bytes(bad)
    with pytest.raises(RsgError, match="RSGKEY01"):
        S.file_info(tmp_path / "magic.rsgk")
    # end mark
    bad = bytearray(raw)
    bad[-1] ^= 0xFF
    (tmp_path / "end.rsgk").write_bytes(bad)
    with pytest.raises(RsgError, match="end mark"):
        S.read_encodings(tmp_path / "end.rsgk", N_R, q, N_E, Q)
    # a non-canonical word is refused on the way out
    w2 = w.copy()
    w2[0, 0] = Q[0]
    with pytest.raises(RsgError, match="canonical"):
        S.write_encodings(tmp_path / "nc.rsgk", w2, N_R, q, N_E, Q)


@pytest.mark.gpu
def test_crs_and_proof_through_hbm(tmp_path):
    """Arena range -> file -> another arena, word for word; a proof produced from the reloaded key equals the proof from the
    original key; a file for other primes is refused by rsg_crs_load."""
    import glob
    import ringsnark_b200 as rs
    from rsgv import Case
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    case = Case(sorted(glob.glob(os.path.join(root, "tests", "golden", "tiny_fast.rsgv")))[0])
    ctx = rs.Context(case.N_R, case.q, case.N_E, case.Q)
    try:
        crs = ctx.crs(5)
        crs.fill_uniform(11)
        S.save_crs(crs, tmp_path / "range.rsgk", first=1, count=3)
        info = S.file_info(tmp_path / "range.rsgk")
        assert info["kind"] == S.FILE_CRS and info["n_elems"] == 3 and info["Q"] == [int(x) for x in case.Q]
        other = ctx.crs(4)
        other.fill_uniform(12)
        assert S.load_crs(other, tmp_path / "range.rsgk", first=1) == 3
        assert np.array_equal(other.download(1, 3), crs.download(1, 3))
        host, kind = S.read_encodings(tmp_path / "range.rsgk", case.N_R, case.q, case.N_E, case.Q)
        assert kind == S.FILE_CRS and np.array_equal(host, crs.download(1, 3))
        with pytest.raises(RsgError, match="too small"):
            S.load_crs(ctx.crs(2), tmp_path / "range.rsgk")
        # a corrupt payload is detected BEFORE anything is written: the arena keeps what it held
        raw = bytearray((tmp_path / "range.rsgk").read_bytes())
        raw[len(raw) // 2] ^= 0x01
        (tmp_path / "flip.rsgk").write_bytes(raw)
        keep = ctx.crs(4)
        keep.fill_uniform(13)
        before = keep.download(0, 4).copy()
        with pytest.raises(RsgError, match="arena untouched|canonical"):
            S.load_crs(keep, tmp_path / "flip.rsgk", first=1)
        assert np.array_equal(keep.download(0, 4), before)
        # a proving key saved and reloaded proves the same proof (= the reference's)
        sys.path.insert(0, os.path.join(root, "tests"))
        r1cs = rs.R1cs(ctx, case.n, case.io, case.aux, case.d["r1cs_row_ptr"], case.d["r1cs_col"], case.d["r1cs_coeff"])
        pk = rs.Groth16ProvingKey(ctx, r1cs)
        pk.load(case.enc("crs_s_pows")[0], case.enc("crs_delta_ts")[0], case.enc("crs_delta_mid")[0], case.enc("crs_alpha")[0],
                case.enc("crs_beta")[0])
        pk.save(tmp_path / "pk.rsgk")
        pk2 = rs.Groth16ProvingKey(ctx, r1cs)
        pk2.load_file(tmp_path / "pk.rsgk")
        assignment = np.concatenate([case.ring("primary_input")[0], case.ring("auxiliary_input")[0]])
        p1, _ = pk.prove(assignment)
        p2, _ = pk2.prove(assignment)
        assert np.array_equal(p1, p2) and np.array_equal(p2, case.enc("proof")[0])
        # the reference's proof survives the container
        proof = case.enc("proof")[0]
        S.write_encodings(tmp_path / "proof.rsgk", proof, case.N_R, case.q, case.N_E, case.Q, kind=S.FILE_PROOF)
        back, kind = S.read_encodings(tmp_path / "proof.rsgk", case.N_R, case.q, case.N_E, case.Q)
        assert kind == S.FILE_PROOF and np.array_equal(back.reshape(-1), np.asarray(proof).reshape(-1))
    finally:
        ctx.close()
