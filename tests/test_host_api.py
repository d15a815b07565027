"""CPU tests of the host-side mirror (no device needed): ingest, region builder, output files, priors."""
import numpy as np
import pytest

import nextgp.jl_b200 as ngp
from nextgp.jl_b200 import api
from oracle import oracle as O


def test_prep_snp_reads_space_delimited_and_drops_missing_columns(tmp_path):
    f = tmp_path / "geno.txt"
    f.write_text("0 1 2 NA\n2 1 0 1\n1 1  2\n".replace("1  2", "1 NA 2"))
    codes = ngp.prep_snp(str(f))
    # column 2 (NA in row 3) and column 3 (NA in row 1) are dropped, prepMatVec.jl:118
    assert codes.dtype == np.int8 and codes.flags.f_contiguous
    assert codes.tolist() == [[0, 1], [2, 1], [1, 1]]


def test_prep_snp_rejects_dosages():
    with pytest.raises(ValueError):
        ngp.prep_snp(np.array([[0.0, 0.5], [1.0, 2.0]]))


@pytest.mark.parametrize("size", [9999, 99, 3, 4, 100])
def test_prep2RegionData_matches_oracle_restatement(tmp_path, size):
    chr_id = np.array([1] * 7 + [2] * 5 + [3] * 3)
    mp = tmp_path / "map.txt"
    mp.write_text("snpID,snpOrder,chrID\n" + "\n".join(f"s{i},{i + 1},{c}" for i, c in enumerate(chr_id)) + "\n")
    offs = ngp.prep2RegionData(str(tmp_path), "M", str(mp), size)
    assert offs.tolist() == O.regions_from_map(chr_id, size).tolist()
    lines = (tmp_path / "groupInfo_M.txt").read_text().splitlines()
    assert lines[0].split("\t") == ["snpID", "snpOrder", "chrID", "groupID"] and len(lines) == 16


def test_outMCMC_and_summaryMCMC_roundtrip(tmp_path):
    d = str(tmp_path)
    ngp.outMCMC(d, "betaM", [["M1", "M2", "M3"]])
    ngp.outMCMC(d, "betaM", np.array([1.0, 2.0, 3.0]))
    ngp.outMCMC(d, "betaM", np.array([3.0, 2.0, 1.0]))
    ngp.outMCMC(d, "varE", [["e"]])
    ngp.outMCMC(d, "varE", 0.25)
    txt = (tmp_path / "betaMOut").read_text().splitlines()
    assert txt[0] == "M1\tM2\tM3" and txt[1] == "1.0\t2.0\t3.0"
    assert ngp.summaryMCMC("betaM", outFolder=d).tolist() == [[2.0, 2.0, 2.0]]
    assert ngp.summaryMCMC("varE", outFolder=d).tolist() == [[0.25]]


def test_prior_constructors_mirror_runTime():
    assert ngp.BayesPR(9999, 0.001).name == "BayesPR" and ngp.BayesPR(9999, 0.001).r == 9999
    b = ngp.BayesC(0.05, 0.01, estimatePi=True)
    assert (b.pi, b.v, b.name, b.estimatePi) == (0.05, 0.01, "BayesC", True)
    assert ngp.BayesB(0.1, 0.02).estimatePi is False
    assert ngp.Random("I", 150.0).v == 150.0


def test_runLMEM_rejects_terms_outside_the_hot_path():
    with pytest.raises(NotImplementedError):
        ngp.runLMEM("y ~ 1 + herd + SNP(M,geno.txt)", {"y": [1.0]}, 10, 2, 2, matrices={"M": np.zeros((1, 1))})


def test_synth_codes_match_oracle_generator():
    pr = ngp.synth.problem(203, 40, 12)
    a = ngp.synth.codes(12, 203, np.arange(40), pr["thr0"], pr["thr1"])
    b = O.synth_codes(12, 203, 0, 40, pr["thr0"], pr["thr1"])
    assert np.array_equal(a, b)
    assert set(np.unique(a)) <= {0, 1, 2}


# ----------------------------------------------------------------------------- native ingest, host-only code of libngp
def _py_parse(path):
    """The reference's rule restated in Python: split on ONE space, "" / NA / NaN / missing are missing, a column with a missing
    value is dropped (prepMatVec.jl:116-118)."""
    rows = []
    with open(path, newline="") as f:
        for line in f.read().split("\n"):
            line = line.rstrip("\r")
            if line == "":
                continue
            rows.append([np.nan if t.strip() == "" or t.strip().upper() in ("NA", "NAN", "MISSING") else float(t) for t in line.split(" ")])
    raw = np.array(rows)
    keep = ~np.isnan(raw).any(axis=0)
    return raw[:, keep].astype(np.int8), keep


@pytest.mark.parametrize("n,p,eol", [(7, 5, "\n"), (64, 33, "\n"), (101, 17, "\r\n")])
def test_native_text_reader_is_bit_exact_and_drops_columns_with_missing(tmp_path, n, p, eol):
    from nextgp.jl_b200 import api
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 3, size=(n, p))
    toks = codes.astype(str).astype(object)
    toks[rng.integers(0, n), 1] = "NA"
    toks[rng.integers(0, n), p - 1] = ""
    toks[0, 3] = "missing"
    toks[n - 1, 0] = "2.0"
    codes[n - 1, 0] = 2
    f = tmp_path / "geno.txt"
    f.write_bytes((eol.join(" ".join(r) for r in toks) + eol).encode())
    packed, n_read, keep = api.read_text_packed(str(f))
    ref, keep_ref = _py_parse(str(f))
    assert n_read == n and np.array_equal(keep, keep_ref) and keep.sum() == p - 3
    assert np.array_equal(api.unpack2(packed, n), ref)
    assert np.array_equal(ngp.prep_snp(str(f)), ref)
    # the padding bits of the last byte of every column are zero (device packing relies on it)
    if n % 4:
        assert not (packed[-1] >> (2 * (n % 4))).any()


@pytest.mark.parametrize("n,p,eol,last_eol", [(7, 5, "\n", True), (257, 33, "\n", False), (1001, 17, "\r\n", True), (513, 40, "\n", True), (256, 9, "\r\n", False)])
def test_native_text_reader_fast_path_matches_the_general_parser(tmp_path, n, p, eol, last_eol):
    """A file of plain 0/1/2 codes takes the threaded block reader (256 rows per task); the same data with one token spelled "2.0" takes
    the general parser: identical packed bytes, zero padding bits, nothing dropped."""
    from nextgp.jl_b200 import api
    rng = np.random.default_rng(n + p)
    codes = rng.integers(0, 3, size=(n, p))
    codes[n - 1, 0] = 2
    toks = codes.astype(str).astype(object)
    f = tmp_path / "clean.txt"
    f.write_bytes((eol.join(" ".join(r) for r in toks) + (eol if last_eol else "")).encode())
    packed, n_read, keep = api.read_text_packed(str(f))
    assert n_read == n and keep.all() and packed.shape == ((n + 3) // 4, p)
    assert np.array_equal(api.unpack2(packed, n), codes.astype(np.int8))
    if n % 4:
        assert not (packed[-1] >> (2 * (n % 4))).any()
    toks[n - 1, 0] = "2.0"
    g = tmp_path / "general.txt"
    g.write_bytes((eol.join(" ".join(r) for r in toks) + (eol if last_eol else "")).encode())
    packed2, n2, keep2 = api.read_text_packed(str(g))
    assert n2 == n and np.array_equal(packed, packed2) and np.array_equal(keep, keep2)


def test_native_text_reader_falls_back_on_anything_but_plain_codes(tmp_path):
    from nextgp.jl_b200 import api
    f = tmp_path / "g.txt"
    f.write_text("0 1 2\n1 3 2\n")                 # right shape, a character outside 0..2: general parser -> not a code
    with pytest.raises(ValueError):
        api.read_text_packed(str(f))
    f.write_text("0 1 2\n0  12\n")                 # right length, wrong structure
    with pytest.raises(ValueError):
        api.read_text_packed(str(f))
    f.write_text("0 1 2\n\n2 1 0\n")               # an empty line between rows is skipped by both paths
    packed, n, keep = api.read_text_packed(str(f))
    assert n == 2 and np.array_equal(api.unpack2(packed, 2), np.array([[0, 1, 2], [2, 1, 0]], dtype=np.int8))


def test_native_text_reader_rejects_dosages_and_ragged_rows(tmp_path):
    from nextgp.jl_b200 import api
    f = tmp_path / "g.txt"
    f.write_text("0 1 2\n1 0.5 2\n")
    with pytest.raises(ValueError):
        api.read_text_packed(str(f))
    f.write_text("0 1 2\n1 0\n")
    with pytest.raises(ValueError):
        api.read_text_packed(str(f))


@pytest.mark.parametrize("n,p", [(8, 3), (13, 40), (1001, 7)])
def test_plink_bed_reader_is_bit_exact(tmp_path, n, p):
    from nextgp.jl_b200 import api
    rng = np.random.default_rng(p)
    a1 = rng.integers(0, 3, size=(n, p))                 # copies of allele A1
    miss = np.zeros((n, p), dtype=bool)
    miss[rng.integers(0, n), 2] = True
    enc = np.where(a1 == 2, 0b00, np.where(a1 == 1, 0b10, 0b11))
    enc = np.where(miss, 0b01, enc)
    bpc = (n + 3) // 4
    bed = bytearray([0x6c, 0x1b, 0x01])
    for j in range(p):
        col = np.zeros(bpc * 4, dtype=np.uint8)
        col[:n] = enc[:, j]
        b = col[0::4] | (col[1::4] << 2) | (col[2::4] << 4) | (col[3::4] << 6)
        bed += bytes(b.astype(np.uint8))
    f = tmp_path / "x.bed"
    f.write_bytes(bytes(bed))
    for count_a1 in (True, False):
        packed, keep = api.read_bed_packed(str(f), n, p, count_a1=count_a1)
        assert keep.sum() == p - 1 and not keep[2]
        want = (a1 if count_a1 else 2 - a1)[:, keep].astype(np.int8)
        assert np.array_equal(api.unpack2(packed, n), want)
    f.write_bytes(bytes([0x6c, 0x1b, 0x00]) + bytes(bed[3:]))
    with pytest.raises(ngp.NgpError):
        api.read_bed_packed(str(f), n, p)               # individual-major files are not supported
