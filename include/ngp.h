/*
 * ngp.h — C ABI of libngp.so: the B200-native (sm_100a) marker-effect Gibbs
 * sweep that slots in behind NextGP.jl's per-marker-set sampler.
 *
 * Every entry point cites the reference interface it replaces (file:line under
 * the NextGP.jl v1.2.0 tree).  Plain pointers and sizes only; every export
 * returns 0 on success and a negative NGP_E* code otherwise, with the text in
 * ngp_last_error().  There is NO CPU backend: without a CUDA device every
 * compute entry point fails with NGP_ECUDA.
 *
 * One handle = one chain on one GPU.  A handle is not thread-safe.
 * Host buffers are borrowed for the duration of a call only.
 */
#ifndef NGP_H_
#define NGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGP_ABI_VERSION 1
#define NGP_MAX_SETS 8

typedef struct ngp_handle ngp_handle;

enum ngp_status {
    NGP_OK = 0,
    NGP_EINVAL = -1,   /* bad argument / call order                         */
    NGP_ECUDA = -2,    /* CUDA runtime error (text in ngp_last_error)        */
    NGP_EDATA = -3,    /* genotype value outside {0,1,2} (missing -> reject; */
                       /* the reference drops such columns, prepMatVec.jl:118)*/
    NGP_ERANGE = -4,   /* fixed-point reduction overflow guard tripped       */
    NGP_ENOMEM = -5,
    NGP_EUNSUPPORTED = -6,
    NGP_ENUMERIC = -7, /* a covariance matrix of the tuple sampler lost positive definiteness */
    NGP_ETIMEOUT = -8  /* row-sharded chain: another rank did not arrive within 4 s (its launch failed or it is not running); results invalid */
};

/* priorVCV[pSet].name dispatch of mme.jl:331,350,362 */
enum ngp_method { NGP_BAYESPR = 0, NGP_BAYESB = 1, NGP_BAYESC = 2,
                  NGP_BAYESR = 3 /* mme.jl:374-383 */,
                  NGP_BAYESRCPI = 5 /* mme.jl:385-403, functions.jl:291-360 */, NGP_BAYESRCPLUS = 6 /* mme.jl:405-418, functions.jl:362-419 */ };
#define NGP_MAX_CLASSES 8

/* host input formats for ngp_upload_genotypes */
enum ngp_geno_format {
    NGP_GENO_I8 = 0,      /* int8 codes 0/1/2, column-major, leading dimension ld        */
    NGP_GENO_F64 = 1,     /* Float64 0.0/1.0/2.0 (raw, NOT centred), column-major, ld     */
                          /*   = the Matrix{Float64} of prepMatVec.jl:120 before :129     */
    NGP_GENO_PACKED2 = 2  /* 2-bit codes, 4 per byte, LSB first, columns padded to bytes; */
                          /*   ld = bytes per column (>= ceil(n/4))                       */
};

/* device storage of the genotype codes */
enum ngp_storage {
    NGP_STORE_I8 = 0,     /* one byte per code, row panels x marker blocks, tiles in INT8-MMA operand order */
    NGP_STORE_2BIT = 1    /* four codes per byte: the same tiles, every 32-bit word (4 rows of a marker) squeezed into one  */
                          /* byte and expanded to the INT8 operands on chip; a quarter of the HBM bytes per sweep.  Blocked   */
                          /* sweep of BayesPR / BayesB / BayesC sets on one GPU (per-marker kernel, tuple: NGP_EUNSUPPORTED); */
                          /* all sets of a handle share the storage format                                                  */
};

/* kernel variants (ngp_configure key NGP_CFG_KERNEL) */
enum ngp_kernel {
    NGP_KERNEL_BLOCKED = 0,  /* blocked exact sweep: one grid-wide reduction per block of markers */
    NGP_KERNEL_LITERAL = 1   /* one grid-wide reduction per marker (north-star baseline design)   */
};

enum ngp_config_key {
    NGP_CFG_KERNEL = 0,        /* enum ngp_kernel                                           */
    NGP_CFG_BLOCK = 1,         /* markers per block: 16, 32 or 64 (0 = auto)                */
    NGP_CFG_MIN_ROWS = 2,      /* minimum rows per worker CTA when choosing the panel count */
    NGP_CFG_MAX_CTAS = 3,      /* cap on CTAs incl. the chain CTA (0 = one per SM)          */
    NGP_CFG_LOOKAHEAD = 4,     /* blocks of look-ahead D of the blocked sweep (0 = auto)    */
    NGP_CFG_TILE_STAGES = 5,   /* genotype tile ring stages per worker CTA, >= D+2 (0 = auto) */
    NGP_CFG_NEAR = 6,          /* cross-Gram distances kept in the chain CTA's block record (0 = auto) */
                               /* keys 1..6 must be set before the first upload             */
    NGP_CFG_PROFILE = 7,       /* 1 = launch the instrumented kernel (cycle counters for ngp_get_profile) */
    NGP_CFG_VERSIONS = 9,      /* versions of the fixed-point residual kept per worker CTA (0 = auto; before the first upload) */
    NGP_CFG_REFETCH = 10,      /* tile ring: 0 tiles stay resident until applied to e, 1 only until their dots are formed (changed columns re-read from L2), -1 auto; before the first upload */
    NGP_CFG_OPT = 11,          /* bit mask of schedule options of the blocked sweep (results unchanged): 1 = genotype tiles streamed with an L2 evict-first policy, 2 = far cross-Gram rows prefetched into L2 when an effect changes; row-sharded blocked sweep: 512 = rank-local pre-reduction off (on by default); -1 = auto */
    NGP_CFG_DEBUG = 8          /* timing experiments that decouple the kernel's roles; RESULTS ARE INVALID when non-zero */
};

/* -------------------------------------------------------------------------
 * Prior of one marker set: what mme.getMME! derives from priorVCV[pSet]
 * (mme.jl:324-373 method wiring, :492-506 df/scale, :513-520 initial varBeta).
 * ------------------------------------------------------------------------- */
typedef struct ngp_prior {
    int32_t method;              /* enum ngp_method                                          */
    int32_t est_pi;              /* BayesB/C: estimatePi (mme.jl:358,370)                     */
    double df;                   /* 3 + size(v,1) = 4 (mme.jl:493)                            */
    double scale;                /* v*(df-2)/df (mme.jl:501)                                  */
    double var_init;             /* v: initial value of every variance slot (mme.jl:516)      */
    double pi_in;                /* BayesB/C: prior inclusion probability (mme.jl:351,363)    */
    int64_t n_regions;           /* BayesPR: length(regionArray) (mme.jl:335-348); B/C: ignored */
    const int64_t* region_off;   /* BayesPR: n_regions+1 offsets, 0-based half-open; NULL = one region */
    const double* lhs0;          /* p or NULL: M[pSet][:lhs] summary-stat precision (mme.jl:314-322) */
    const double* rhs0;          /* p or NULL: M[pSet][:rhs]                                   */
    /* BayesR only (runTime.jl:78-93, mme.jl:374-383)                                            */
    int32_t n_class;             /* length(class) <= NGP_MAX_CLASSES                           */
    int32_t pad_;
    const double* v_class;       /* [n_class] M[pSet][:vClass], e.g. 0, 1e-4, 1e-3, 1e-2        */
    const double* pi_class;      /* [n_class] starting / fixed class proportions (est_pi: Dirichlet update, functions.jl:284-288) */
} ngp_prior;

/* -------------------------------------------------------------------------
 * Prior of a TUPLE of marker sets (multi-breed / correlated effects): what
 * mme.getMME! derives for `(:M1,:M2) => BayesPR(r, V)` (mme.jl:448-489 layout,
 * :493 df = 3 + k, :501 scale = V .* (df - k - 1), :516 initial varBeta = V).
 * The member sets share n, p and the region structure; per locus their k effects
 * are drawn jointly (functions.jl:140-154).  Row-major k x k matrices.
 * ------------------------------------------------------------------------- */
typedef struct ngp_joint_prior {
    int32_t k;                       /* 2..8 member sets                                         */
    int32_t set_id[NGP_MAX_SETS];    /* uploaded marker sets; breed b = set_id[b] (the tuple replaces them, mme.jl:459) */
    int32_t pad_;
    double df;                       /* 3 + k                                                    */
    const double* scale;             /* [k*k]                                                    */
    const double* var_init;          /* [k*k] V: initial value of every region's covariance      */
    int64_t n_regions;               /* length(regionArray) (mme.jl:470-487)                     */
    const int64_t* region_off;       /* n_regions+1 offsets, 0-based half-open; NULL = one region */
} ngp_joint_prior;

/* -------------------------------------------------------------------------
 * Row-sharded single chain (BASELINE config 5): rank r of `world`
 * holds a contiguous slice of the individuals (rows of X, y, e).  Per marker the
 * partial dots of all ranks are reduced through NVLink peer memory inside the
 * persistent kernel; every rank draws the same effect from the same counter.
 * What a rank publishes to its peers (exchange the structs with any transport,
 * e.g. an all-gather over NCCL/gloo, or pass them by hand within one process):
 * ------------------------------------------------------------------------- */
typedef struct ngp_shard_info {
    unsigned char ipc[64];   /* cudaIpcMemHandle_t of the rank's synchronisation area            */
    int64_t n_local;         /* rows held by the rank                                            */
    int32_t worker_ctas;     /* row panels (worker CTAs) of the rank                             */
    int32_t device;
    int64_t pid;             /* same pid => same-process attach through local_ptr (+ peer access) */
    uint64_t local_ptr;
} ngp_shard_info;

/* -------------------------------------------------------------------------
 * One fixed-effect term besides the intercept: X[xSet] of getMME! (covariate,
 * or the level columns of a factor).  Sampled after the intercept, in the order
 * given, by sampleX! (functions.jl:39-54): a single column like the intercept,
 * several columns with "Wang's trick" (sampleb!, functions.jl:22-36).
 * ------------------------------------------------------------------------- */
typedef struct ngp_fixed_set {
    int32_t n_cols;              /* columns of X[xSet].data                                   */
    int32_t pad_;
    const double* data;          /* [n x n_cols] column-major                                  */
    double lhs0, rhs0;           /* single-column sets: X[xSet].lhs / .rhs (else ignored)      */
} ngp_fixed_set;
#define NGP_MAX_FIXED_SETS 4
#define NGP_MAX_FIXED_COLS 32

/* Variate log of n_iter iterations for replay parity: the sampler
 * consumes these instead of its Philox stream.  Layout is row-major
 * [iteration][index].  Set s uses u[s], z[s], chi2_b[s], beta_pi[s].           */
typedef struct ngp_replay {
    int32_t n_iter;
    int32_t n_sets;
    const double* chi2_e;                  /* [n_iter]       functions.jl:524               */
    const double* z_mu;                    /* [n_iter]       functions.jl:45                */
    const double* u[NGP_MAX_SETS];         /* [n_iter][p]    functions.jl:174,216 (B/C); BayesR: [n_iter][p][n_class], one per comparison (functions.jl:261) */
    const double* z[NGP_MAX_SETS];         /* [n_iter][p]    functions.jl:494               */
    const double* chi2_b[NGP_MAX_SETS];    /* [n_iter][nvar] functions.jl:510 (nvar: PR=n_regions, B=p, C=1) */
    const double* beta_pi[NGP_MAX_SETS];   /* [n_iter]       functions.jl:532; BayesR: [n_iter][n_class] the Dirichlet draw (functions.jl:537) */
} ngp_replay;

/* Chain state, the arguments M[mSet].funct mutates (functions.jl:118,157,197)
 * plus what runSampler! keeps between iterations (samplers.jl:23).  NULL
 * members are skipped.                                                          */
typedef struct ngp_state {
    int64_t n;                             /* individuals                                    */
    int32_t n_sets;
    int32_t pad_;
    double* e;                             /* [n] ycorr                                      */
    double mu;                             /* b[intercept]                                   */
    double varE;
    int64_t iter;                          /* iterations completed on this handle            */
    double* beta[NGP_MAX_SETS];            /* [p]                                            */
    int64_t* delta[NGP_MAX_SETS];          /* [p] Int64 like mme.jl:444                      */
    double* varBeta[NGP_MAX_SETS];         /* [nvar]                                         */
    double pi[NGP_MAX_SETS][2];            /* piHat = [not fitted, fitted] (mme.jl:359,371)   */
} ngp_state;

typedef struct ngp_timing {
    double last_run_ms;          /* device time of the last ngp_run / ngp_sweep kernel (CUDA events) */
    int64_t launches;            /* kernels launched by this handle since creation              */
    int32_t ctas, threads;       /* geometry of the sweep kernel                                */
    int32_t block, rows_per_cta; /* markers per block, rows per CTA panel                        */
    int64_t smem_bytes;
    int32_t lookahead, near_depth, tile_stages, record_stages;
    int32_t kernel_variant;      /* sweep kernel of the last launch: 0 blocked, 1 instrumented, 2 timing experiments, 3 per-marker (chosen by the library for */
                                 /* weighted residuals, summary-statistic BayesR, > 4 BayesR classes, row sharding), 4 tuple, 5 shard group, 6 blocked     */
                                 /* with > 512 rows per CTA, 7 blocked BayesR; -1 none yet                                                                */
    int32_t refetch, storage_2bit, pad_;   /* tile-ring mode, device storage of the genotypes */
    double sum_run_ms;           /* device time of ALL ngp_run / ngp_sweep kernels of this handle so far (sum of the per-launch event times) */
} ngp_timing;

/* ---- lifetime ------------------------------------------------------------ */
int ngp_abi_version(void);
int ngp_device_count(void);
int ngp_create(int device, ngp_handle** out);
int ngp_destroy(ngp_handle* h);
const char* ngp_last_error(const ngp_handle* h);   /* h may be NULL: error of the last failed ngp_create */
int ngp_configure(ngp_handle* h, int key, int64_t value);
/* run on a caller-owned CUDA stream (cudaStream_t as void*); NULL = handle's own stream */
int ngp_set_stream(ngp_handle* h, void* cuda_stream);

/* ---- data: replaces prepMatVec.prep's SNP branch output M[arg1][:data]
 *      (prepMatVec.jl:116-131) and getMME!'s mpm/Mp (mme.jl:299-311).
 *      Packs on device, computes column sums, means, mpm and block Gram.   -- */
int ngp_upload_genotypes(ngp_handle* h, int set_id, int64_t n, int64_t p,
                         const void* data, int fmt, int64_t ld, int storage);
/* synthetic codes generated on device: code(i,j) =
 * (w>=thr0[j]) + (w>=thr1[j]) with w = word (i&3) of Philox4x32-10(key=seed,
 * ctr=(i>>2, j, 0, 0x47454e4f)).  Bit-identical to oracle ngo_synth_codes.  */
int ngp_synth_genotypes(ngp_handle* h, int set_id, int64_t n, int64_t p, uint64_t seed,
                        const uint32_t* thr0, const uint32_t* thr1, int storage);
/* rows [row0, row0 + n) of the same synthetic matrix (row0 a multiple of 4): the slice of one rank of a row-sharded chain */
int ngp_synth_genotypes_rows(ngp_handle* h, int set_id, int64_t row0, int64_t n, int64_t p, uint64_t seed,
                             const uint32_t* thr0, const uint32_t* thr1, int storage);
/* unpack device storage back to int8 column-major (bit-exact round-trip tests) */
int ngp_download_genotypes(ngp_handle* h, int set_id, int64_t j0, int64_t j1, int8_t* out);
/* mean_j (prepMatVec.jl:129) and mpm_j = X_j'X_j of the centred column (mme.jl:305-307) */
int ngp_get_column_stats(ngp_handle* h, int set_id, double* mean, double* mpm);

/* ---- row-sharded chain: call order  ngp_shard_init -> uploads (local rows) -> ngp_shard_export -> [exchange] ->
 *      ngp_shard_attach -> ngp_get_column_sums -> [all-reduce] -> ngp_set_column_sums -> priors, phenotype (local rows) ->
 *      ngp_run with identical arguments on every rank, concurrently.  mean_j and mpm_j (prepMatVec.jl:129, mme.jl:305-307)
 *      are those of the whole column.  Per-marker kernel (NGP_KERNEL_LITERAL), or the blocked kernel after ngp_set_gram. */
int ngp_shard_init(ngp_handle* h, int rank, int world);
int ngp_shard_export(ngp_handle* h, ngp_shard_info* out);
int ngp_shard_attach(ngp_handle* h, const ngp_shard_info* all_ranks);
/* All ranks of a chain whose shards live on ONE device (fewer GPUs than ranks: tests, small boxes): one cooperative grid over all
 * ranks' row slices, handles[r] = rank r.  Kernels that wait for one another must never be separate launches on one GPU (they are
 * not guaranteed to be co-resident), so ngp_run refuses such a handle and this call replaces it.  Same protocol and code path as
 * between GPUs (system-scope REDs / polls on every rank's synchronisation area).  A rank that does not arrive within 4 s ends every
 * wait: NGP_ETIMEOUT.                                                                                                          */
int ngp_run_group(ngp_handle** handles, int n_handles, int32_t n_iter);
int ngp_get_column_sums(ngp_handle* h, int set_id, int64_t* colsum, int64_t* colsumsq);
int ngp_set_column_sums(ngp_handle* h, int set_id, int64_t n_total, const int64_t* colsum, const int64_t* colsumsq);
/* Row-sharded chain on the BLOCKED kernel (NGP_KERNEL_BLOCKED; SURVEY 8e "B-many scalars blocked"): per block every worker CTA pushes its B
 * partial sums into the accumulator ring of every rank, each rank's chain CTA computes the identical lists (identical sums, identical
 * draws).  The banded Gram is a sum over individuals: read every rank's (ngp_get_gram), add them up, write the total back on every rank
 * (ngp_set_gram) — all ranks must share the block size and the look-ahead.  count = (p_pad / B) (D + 1) B B int32 values.            */
int ngp_gram_size(ngp_handle* h, int set_id, int64_t* count);
int ngp_get_gram(ngp_handle* h, int set_id, int32_t* out);
int ngp_set_gram(ngp_handle* h, int set_id, const int32_t* in);

/* host 2-bit codec (format NGP_GENO_PACKED2); returns NGP_EDATA on a code outside 0..2 */
int ngp_pack2(const int8_t* codes, int64_t n, int64_t p, int64_t ld_in, uint8_t* out, int64_t ld_out);
int ngp_unpack2(const uint8_t* packed, int64_t n, int64_t p, int64_t ld_in, int8_t* out, int64_t ld_out);

/* ---- ingest straight to 2-bit codes (never through the Matrix{Float64} of prepMatVec.jl:116-120).  Host-only.
 * Text: the reference's format — one row per individual, fields separated by ONE space, no header; a column with a missing
 * field ("", NA, NaN, missing) is dropped like prepMatVec.jl:118 does; any value other than 0/1/2 -> NGP_EDATA.
 * Call once with packed == NULL to learn n and p_total, allocate ld * p_total bytes (ld >= ceil(n/4)) and keep[p_total], call again:
 * the kept columns are compacted to the front (NGP_GENO_PACKED2 layout), keep[j] tells which survived, *p_kept how many.        */
int ngp_read_text_genotypes(const char* path, int64_t* n, int64_t* p_total, uint8_t* packed, int64_t ld, uint8_t* keep, int64_t* p_kept);
/* PLINK .bed (SNP-major): n samples, p variants; code = copies of allele A1 (count_a1 != 0) or A2.  Variants with a missing
 * genotype are dropped the same way.                                                                                          */
int ngp_read_bed_genotypes(const char* path, int64_t n, int64_t p, int count_a1, uint8_t* packed, int64_t ld, uint8_t* keep, int64_t* p_kept);

/* ---- model: replaces the state getMME! allocates (mme.jl:57,87-94,443-444,492-520) */
int ngp_set_phenotype(ngp_handle* h, const double* y, int64_t n);            /* ycorr = deepcopy(Y) */
int ngp_set_residual_prior(ngp_handle* h, double df_e, double scale_e);      /* E[:df], E[:scale]  */
/* E.str == "D" (mme.jl:70-73, 133-136, 299-303; functions.jl:526-528): w = E.iVarStr = inv.(D), n positive weights; NULL = "I".
 * Fixed effects + marker-set models (BayesPR/B/C/R) on one GPU; sampled by the per-marker kernel.                            */
int ngp_set_residual_weights(ngp_handle* h, const double* w, int64_t n);
/* single-column fixed effect of ones (functions.jl:39-47); lhs0/rhs0 = X[xSet][:lhs/:rhs] */
int ngp_set_intercept(ngp_handle* h, int enabled, double lhs0, double rhs0);
/* fixed effects besides the intercept (at most NGP_MAX_FIXED_SETS sets, NGP_MAX_FIXED_COLS columns in all); call after the first
 * upload; n_sets = 0 removes them.  Sampled inside ngp_run right after the intercept (samplers.jl:37-39).           */
int ngp_set_fixed_effects(ngp_handle* h, int n_sets, const ngp_fixed_set* sets);
int ngp_get_fixed_effects(ngp_handle* h, double* b);                          /* all columns, set after set */
/* replay: one standard normal per column and iteration, z[n_iter][n_cols] (functions.jl:34,45); after ngp_set_replay */
int ngp_set_fixed_replay(ngp_handle* h, int32_t n_iter, const double* z);
int ngp_set_prior(ngp_handle* h, int set_id, const ngp_prior* prior);
/* Replace only lhs0 / rhs0 (p doubles each, or NULL = none) of a set that already has a prior; coefficients, indicators and posterior
 * sums are kept.  The GRN single-site loop (GRN.jl:150-164, sampleΛ2!) runs the same markers once per gene with the offset
 * alpha*pMeans[g] on the right-hand side: see INTEGRATION.md "GRN". */
int ngp_set_marker_summary(ngp_handle* h, int set_id, const double* lhs0, const double* rhs0);
/* Overwrite varBeta (nvar doubles: regions for BayesPR, p for BayesB, 1 else) of one set, nothing else.  BayesLV (functions.jl:421-486)
 * = the BayesPR sweep with one region per locus, followed by the caller's own model of the log-variances (functions.jl:446-485), which
 * replaces the scaled-inverse-chi-square draw of the device: see INTEGRATION.md "BayesLV". */
int ngp_set_var_beta(ngp_handle* h, int set_id, const double* varBeta);
/* tuple of marker sets with jointly drawn effects (mme.jl:448-489); at most one tuple per handle, and every
 * uploaded set of the handle must be a member                                                            */
int ngp_set_joint_prior(ngp_handle* h, const ngp_joint_prior* prior);

/* ---- variates --------------------------------------------------------------- */
int ngp_set_rng(ngp_handle* h, uint64_t seed, uint32_t chain_id);            /* Philox4x32-10 key / stream */
int ngp_set_replay(ngp_handle* h, const ngp_replay* log);                    /* NULL = back to Philox */
/* variate log of the tuple: z [n_iter][p][k] (MvNormal = mean + chol(C) z, functions.jl:149), Bartlett variates of the
 * inverse-Wishart draw (functions.jl:515) iw_chi2 [n_iter][n_regions][k], iw_z [n_iter][n_regions][k][k] (strict lower
 * triangle).  Call after ngp_set_replay (which carries chi2_e / z_mu; its per-set members are ignored for tuple members). */
int ngp_set_joint_replay(ngp_handle* h, int32_t n_iter, const double* z, const double* iw_chi2, const double* iw_z);

/* ---- sampling ---------------------------------------------------------------
 * ngp_run replaces n_iter trips of the loop body of samplers.runSampler!
 * (samplers.jl:29-53): varE -> intercept -> every marker set, all on device. */
int ngp_run(ngp_handle* h, int32_t n_iter);
/* ngp_sweep replaces ONE call M[mSet].funct(mSet,M,beta,delta,ycorr,varE,varBeta)
 * (samplers.jl:52; functions.jl:118,157,197): host buffers in, mutated in place. */
int ngp_sweep(ngp_handle* h, int set_id, double* ycorr, double varE,
              double* beta, int64_t* delta, double* varBeta, double* piHat);   /* piHat: 2 values (B/C) or n_class (BayesR) */
/* ngp_joint_sweep replaces ONE call sampleBayesPR!(mSet::Tuple, M, beta, delta, ycorr, varE, varBeta)
 * (functions.jl:140-154): beta is k x p row-major (row b = breed b), varBeta is n_regions x k x k.          */
int ngp_joint_sweep(ngp_handle* h, double* ycorr, double varE, double* beta, double* varBeta);
/* effects (k x p row-major) and region covariances (n_regions x k x k) of the tuple; NULL members are skipped */
int ngp_get_joint_state(ngp_handle* h, double* beta, double* varBeta);
/* BayesR: class proportions M[pSet][:piHat] (n_class values); ngp_state.delta holds the 1-based class of every locus */
int ngp_get_class_pi(ngp_handle* h, int set_id, double* piHat);
/* ---- BayesRCpi / BayesRCplus: priorVCV[pSet] = BayesRCπ(pi, class, v, annot) / BayesRCplus(...) (runTime.jl:95-113); what getMME!
 *      derives (mme.jl:385-418): one variance and one vector of class proportions PER ANNOTATION (varBeta has n_annot entries), annotProb =
 *      annot ./ rowsums, annotCat.  Swept by the per-marker kernel; n_annot * n_class <= 32 (one lane per annotation and class).
 *      ngp_state.delta = 1-based class of every locus, ngp_state.varBeta = the n_annot variances.                                          */
typedef struct ngp_rc_prior {
    int32_t plus;                /* 0 = BayesRCpi (sampleBayesRCπ!), 1 = BayesRCplus (sampleBayesRCplus!)                        */
    int32_t est_pi;              /* estimatePi: Dirichlet update of every annotation's class proportions (functions.jl:352-359)  */
    double df, scale, var_init;  /* as ngp_prior (mme.jl:493,501,516)                                                            */
    int32_t n_class, n_annot;
    const double* v_class;       /* [n_class]  M[pSet][:vClass]                                                                  */
    const double* pi_class;      /* [n_class]  priorVCV[pSet].pi: the starting proportions of every annotation (mme.jl:390-392)  */
    const int32_t* annot;        /* [p][n_annot] row-major: M[pSet][:annotInput] (mme.jl:394)                                    */
} ngp_rc_prior;
int ngp_set_rc_prior(ngp_handle* h, int set_id, const ngp_rc_prior* pr);
/* annotCat (1-based, RCpi; p), annotProb (p x n_annot), piHat (n_annot x n_class); NULL members are skipped */
int ngp_get_rc_state(ngp_handle* h, int set_id, int64_t* annot_cat, double* annot_prob, double* pi_hat);
/* replay log of an RC set (after ngp_set_replay, which carries chi2_e / z_mu): u_annot [iter][p] and dirp [iter][p][n_annot] (RCpi: the
 * uniform of the Categorical draw and the RESULT of sampleProb), u [iter][p][n_class] (RCpi) or [iter][p][n_annot][n_class] (RCplus),
 * z [iter][p] or [iter][p][n_annot], chi2_b [iter][n_annot], dir_pi [iter][n_annot][n_class] (NULL unless est_pi)                     */
int ngp_set_rc_replay(ngp_handle* h, int set_id, int32_t n_iter, const double* u_annot, const double* dirp, const double* u,
                      const double* z, const double* chi2_b, const double* dir_pi);
int ngp_get_state(ngp_handle* h, ngp_state* out);
int ngp_set_state(ngp_handle* h, const ngp_state* in);
/* running posterior sums since the last reset: sum(beta), sum(beta^2), sum(delta) per marker */
int ngp_reset_posterior(ngp_handle* h);
int ngp_get_posterior(ngp_handle* h, int set_id, int64_t* n_samples, double* sum_beta, double* sum_beta2, double* sum_delta);
int ngp_get_timing(ngp_handle* h, ngp_timing* out);

/* per-CTA cycle counters of the last launch, 32 int64 per CTA (clock64; blocked kernel).  The last CTA is the
 * chain CTA, the others are worker CTAs.
 * worker CTA (thread 0):  [1] IMMA dots + limb combine [2] axpy + re-quantise [10] wait for the changed-effect list
 *                         [14] wait for the TMA tile [15] CTA combine + RED
 * chain CTA  (thread 0):  [0] wait for the block record [3] wait for r_base (prep warps) [4] scalar chain
 *                         [6] markers whose effect changed [7] speculative evaluations
 *            (first prep warp): [12] far corrections [13] accumulator poll
 * every CTA  (thread 0):  [8] phase 0 (varE, intercept) [9] phase 1 (marker constants) [11] phase 3
 * Returns the number of CTAs written. */
int ngp_get_profile(ngp_handle* h, int64_t* out, int32_t max_ctas);

/* instrumented kernel only: (start clock, cycles waited for the dots) of the chain warp's first 2048 steps of the last sweep */
int ngp_get_trace(ngp_handle* h, int64_t* out, int32_t n);

/* stream self-test: fills out[0..n) with the handle's variates of one purpose
 * (purpose: 2=uniform 3=normal 4=chisq(df)) for iteration iter, set set_id.   */
int ngp_debug_variates(ngp_handle* h, int set_id, uint32_t iter, int purpose, double df, int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* NGP_H_ */
