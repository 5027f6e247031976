// CUDA backend (sm_100a) of the aeonflux_b200 batch engine: the __global__ wrappers around the stage bodies of
// engine.cuh, their launch geometry, and the C ABI (api_impl.inc).
//
// Geometry: every stage is "one thread per (item, job)" with items along x, so a warp always runs one job type on 32
// consecutive items: control flow is warp-uniform (all data-dependent choices are public digits) and every global
// access is a 128-bit load/store of an item-contiguous record.  Work is integer-multiply bound (IMAD.WIDE on the fma
// pipe); there is no GEMM-shaped step, so no tensor cores, and HBM traffic is ~20 KB per presentation.
#include <cuda_runtime.h>

#include <cstdio>

#include "engine.cuh"

namespace afx {

constexpr int TPB = 128;            // threads per CTA for the ladder kernels (4 warps, one per SMSP)
constexpr int TPB_MSM = 512;        // one lockstep CTA per SM for the constraint ladders (16 warps share the I-cache)
constexpr int MSM_STAGE_TABLES = 3; // constant-base tables staged in shared memory per CTA (12 KB each)

__global__ void __launch_bounds__(256) k_scalar_check(Workspace ws, const u16* fields) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) scalar_check_job(ws, fields[blockIdx.y], item);
}

__global__ void __launch_bounds__(TPB, 4) k_points(Workspace ws, const PointJob* jobs) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) points_job(ws, jobs[blockIdx.y], item);
}

__global__ void __launch_bounds__(TPB, 4) k_amac(Workspace ws, const AmacDesc* d) {
    extern __shared__ u32 smem[];
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) amac_job(ws, *d, item, smem + threadIdx.x, blockDim.x);
}

__global__ void __launch_bounds__(TPB_MSM, 1) k_msm(Workspace ws, const MsmDesc* msms, const u32* group_idx, u32 scratch_terms) {
    extern __shared__ __align__(16) u32 smem[];
    const MsmDesc& d = msms[group_idx[blockIdx.y]];
    u32* scratch = smem;                                          // [scratch_terms*8][TPB_MSM]
    u32* staged = smem + (size_t)scratch_terms * 8 * blockDim.x;  // [<=MSM_STAGE_TABLES][128][24]
    u32 nstage = d.ncon < MSM_STAGE_TABLES ? d.ncon : MSM_STAGE_TABLES;
    for (u32 k = 0; k < nstage; k++) {
        const uint4* src = reinterpret_cast<const uint4*>(ws.ctabs + (size_t)d.con[k].ctab * CTAB_ENTRIES * 24);
        uint4* dst = reinterpret_cast<uint4*>(staged + (size_t)k * CTAB_ENTRIES * 24);
        for (u32 i = threadIdx.x; i < CTAB_ENTRIES * 24 / 4; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = item < ws.count;
    if (!active) item = ws.count - 1;                              // keep every thread in the barriers of the ladder
    CtabResolver ctab_of{staged, ws.ctabs, &d, nstage};
    msm_job(ws, d, item, scratch + threadIdx.x, blockDim.x, ctab_of, active);
}

// Issuer::issue: the same per-(item, job) ladders with every lookup constant-address (all scalars are secrets)
__global__ void __launch_bounds__(TPB_MSM, 1) k_msm_ct(Workspace ws, const MsmDesc* msms, const u32* group_idx) {
    extern __shared__ __align__(16) u32 smem[];
    const MsmDesc& d = msms[group_idx[blockIdx.y]];
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = item < ws.count;
    if (!active) item = ws.count - 1;
    msm_ct_job(ws, d, item, smem + threadIdx.x, blockDim.x, active);
}

__global__ void __launch_bounds__(256) k_derive(Workspace ws, const WideDesc* d) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) derive_job(ws, d[blockIdx.y], blockIdx.y, item);
}

__global__ void __launch_bounds__(256) k_issue_out(Workspace ws, const IssueOutDesc* d, u32* out) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) issue_out_job(ws, *d, blockIdx.y, item, out);
}

__global__ void __launch_bounds__(TPB) k_transcript(Workspace ws, const TxDesc* txs) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) transcript_job(ws, txs[blockIdx.y], item);
}

__global__ void __launch_bounds__(256) k_verdict(Workspace ws, uint8_t* verdicts) {
    u32 item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item < ws.count) verdicts[item] = ws.status[item] != 0;
}

// per-issuer setup: one thread per (constant point, multiple)
__global__ void __launch_bounds__(128) k_ctab_setup(const u32* enc, u32 ncp, u32* ctabs, u32* encneg, u32* bad) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncp * CTAB_ENTRIES) return;
    u32 b = t / CTAB_ENTRIES, m = t % CTAB_ENTRIES + 1;
    u32 ok = ctab_entry_job(enc + 8 * b, m, ctabs + ((size_t)b * CTAB_ENTRIES + (m - 1)) * 24);
    if (m == 1) {
        ge p; u32 w[8];
        ge_decompress(p, enc + 8 * b);
        ge_compress(w, ge_neg(p));
        for (int i = 0; i < 8; i++) encneg[8 * b + i] = w[i];
        if (!ok) atomicOr(bad, 1u);
    }
}
__global__ void k_secret_setup(const u32* secsc, u32 nsec, u32* secdig, const u32* Wenc, u32* W_pniels, u32* bad) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nsec) {
        sc s = sc_from_words(secsc + 8 * t);
        if (!sc_is_canonical(s)) atomicOr(bad, 2u);
        u32 rec[8]; sc_recode16(rec, s);
        for (int i = 0; i < 8; i++) secdig[8 * t + i] = rec[i];
    } else if (t == nsec) {
        ge p;
        if (!ge_decompress(p, Wenc)) atomicOr(bad, 1u);
        store_pniels(W_pniels, ge_to_pniels(p));
    }
}

}  // namespace afx

// ---- backend primitives for api_impl.inc -------------------------------------------------------------------------------
using namespace afx;
typedef cudaStream_t be_stream;
#define AFX_BACKEND_NAME "cuda sm_100a"

static int be_set_device(int d) { return cudaSetDevice(d) != cudaSuccess; }
static int be_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 16) != cudaSuccess; }
static void be_free(void* p) { cudaFree(p); }
static int be_h2d(void* d, const void* h, size_t n, be_stream s) { return cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) != cudaSuccess; }
static int be_d2h(void* h, const void* d, size_t n, be_stream s) { return cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) != cudaSuccess; }
static int be_memset(void* d, int v, size_t n, be_stream s) { return cudaMemsetAsync(d, v, n, s) != cudaSuccess; }
static int be_sync(be_stream s) { return cudaStreamSynchronize(s) != cudaSuccess; }
static int be_check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { std::fprintf(stderr, "aeonflux_b200: CUDA launch failed: %s\n", cudaGetErrorString(e)); return -3; }
    return 0;
}
typedef cudaEvent_t be_event;
static void be_event_create(be_event* e) { cudaEventCreate(e); }
static void be_event_destroy(be_event e) { cudaEventDestroy(e); }
static void be_event_record(be_event e, be_stream s) { cudaEventRecord(e, s); }
static int be_event_sync(be_event e) { return cudaEventSynchronize(e) != cudaSuccess; }
static float be_event_elapsed(be_event a, be_event b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
static dim3 grid_for(u32 count, u32 tpb, u32 y) { return dim3((count + tpb - 1) / tpb, y, 1); }

static void be_launch_scalar_check(const Workspace& ws, const u16* d_fields, u32 nf, be_stream s) {
    k_scalar_check<<<grid_for(ws.count, 256, nf), 256, 0, s>>>(ws, d_fields);
}
static void be_launch_points(const Workspace& ws, const PointJob* d_jobs, u32 njobs, be_stream s) {
    k_points<<<grid_for(ws.count, TPB, njobs), TPB, 0, s>>>(ws, d_jobs);
}
static void be_launch_amac(const Workspace& ws, const AmacDesc* d, u32 nps, be_stream s) {
    size_t smem = (size_t)(nps ? nps : 1) * 8 * TPB * 4;
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_amac, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_set = true; }
    k_amac<<<grid_for(ws.count, TPB, 1), TPB, smem, s>>>(ws, d);
}
static void be_launch_msm(const Workspace& ws, const MsmDesc* d_msms, const u32* d_idx, u32 nidx, u32 max_terms, u32 max_con, be_stream s) {
    u32 nstage = max_con < (u32)MSM_STAGE_TABLES ? max_con : (u32)MSM_STAGE_TABLES;
    // largest CTA (<= TPB_MSM threads) whose digit scratch + staged tables fit in shared memory
    u32 tpb = TPB_MSM;
    size_t smem = 0;
    for (;; tpb /= 2) {
        smem = (size_t)max_terms * 8 * tpb * 4 + (size_t)nstage * CTAB_ENTRIES * 96;
        if (smem <= 200 * 1024 || tpb <= 32) break;
    }
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_msm, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr_set = true; }
    k_msm<<<grid_for(ws.count, tpb, nidx), tpb, smem, s>>>(ws, d_msms, d_idx, max_terms);
}
static void be_launch_msm_ct(const Workspace& ws, const MsmDesc* d_msms, const u32* d_idx, u32 nidx, u32 max_terms, be_stream s) {
    u32 tpb = TPB_MSM;
    size_t smem = 0;
    for (;; tpb /= 2) {
        smem = (size_t)max_terms * 8 * tpb * 4;
        if (smem <= 200 * 1024 || tpb <= 32) break;
    }
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_msm_ct, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr_set = true; }
    k_msm_ct<<<grid_for(ws.count, tpb, nidx), tpb, smem, s>>>(ws, d_msms, d_idx);
}
static void be_launch_derive(const Workspace& ws, const WideDesc* d, u32 nd, be_stream s) {
    k_derive<<<grid_for(ws.count, 256, nd), 256, 0, s>>>(ws, d);
}
static void be_launch_issue_out(const Workspace& ws, const IssueOutDesc* d, u32 nwords, u32* out, be_stream s) {
    k_issue_out<<<grid_for(ws.count, 256, nwords), 256, 0, s>>>(ws, d, out);
}
static void be_launch_transcript(const Workspace& ws, const TxDesc* d_txs, u32 ntx, be_stream s) {
    k_transcript<<<grid_for(ws.count, TPB, ntx), TPB, 0, s>>>(ws, d_txs);
}
static void be_launch_verdict(const Workspace& ws, uint8_t* verdicts, be_stream s) {
    k_verdict<<<grid_for(ws.count, 256, 1), 256, 0, s>>>(ws, verdicts);
}
static void be_launch_ctab_setup(const u32* d_enc, u32 ncp, u32* d_ctabs, u32* d_encneg, u32* d_bad, be_stream s) {
    u32 total = ncp * CTAB_ENTRIES;
    k_ctab_setup<<<(total + 127) / 128, 128, 0, s>>>(d_enc, ncp, d_ctabs, d_encneg, d_bad);
}
static void be_launch_secret_setup(const u32* d_secsc, u32 nsec, u32* d_secdig, const u32* d_Wenc, u32* d_W, u32* d_bad, be_stream s) {
    k_secret_setup<<<1, 64, 0, s>>>(d_secsc, nsec, d_secdig, d_Wenc, d_W, d_bad);
}

#include "api_impl.inc"
