"""nextgp.jl_b200 — B200-native (sm_100a) marker-effect Gibbs sweep behind NextGP.jl's sampler slot.

Holds only what the hot path needs: csrc/ (CUDA kernels + the C ABI of include/ngp.h, built in-tree
to libngp.so) and the host-side mirror of the reference interface (api.py).  No CPU fallback."""
from . import _lib
from ._lib import (BAYESB, BAYESC, BAYESPR, GENO_F64, GENO_I8, GENO_PACKED2, KERNEL_BLOCKED, KERNEL_LITERAL, NgpError, build)
from .api import (BayesB, BayesC, BayesLV, BayesPR, BayesR, BayesRCpi, BayesRCplus, MarkerTerm, Random, Sampler, ShardedChain, SummaryStatistics, getMME, outMCMC, prep2RegionData,
                  prep_snp, runLMEM, runSampler, sampleBayesLV, sampleLambda2, summaryMCMC, LogVarModel)
from . import synth

__all__ = ["Sampler", "ShardedChain", "runLMEM", "getMME", "runSampler", "BayesPR", "BayesB", "BayesC", "BayesR", "BayesRCpi", "BayesRCplus", "BayesLV", "LogVarModel", "sampleBayesLV", "Random", "SummaryStatistics",
           "outMCMC", "summaryMCMC", "sampleLambda2", "prep_snp", "prep2RegionData", "MarkerTerm", "synth", "build", "NgpError"]
