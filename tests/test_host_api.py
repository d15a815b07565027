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
