// foldcomp_b200/csrc/fcz_engine.cu -- B200 (sm_100a) FCZ encode/decode engine behind include/fcz_engine.h.
//
// Execution model:
//   * one CTA works on one chain at a time; the chain's coordinates (encode) or blob (decode) are
//     pulled from HBM into shared memory with ONE bulk async copy (cp.async.bulk + mbarrier, the
//     1-D TMA path) and every intermediate array (atom offsets, angles, sin/cos tables, the output
//     blob / coordinates) lives in shared memory; results leave with coalesced 128-bit stores;
//   * chains are binned by size into TIERS on the device (k_*_plan); each tier is one persistent
//     launch (grid = SMs x CTAs/SM for that tier's shared-memory footprint) whose CTAs pull chains
//     from the tier's list through an atomic ticket;
//   * output offsets are exclusive scans over per-chain sizes (k_scan_*), so blobs and coordinates
//     are tightly packed in chain order regardless of the order CTAs finish;
//   * the per-chain algorithm itself is fcz_codec.h (shared with the CPU model used by the tests).
// No tensor cores: there is no dense contraction on this path.
// Compile with -fmad=false: the encode arithmetic must not be contracted into FMAs.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "fcz_codec.h"
#include "fcz_text.h"
#include "fcz_parse.h"

using namespace fcz;

// ============================================================================================ tiers

struct TierCfg {
    uint32_t max_res;    // residues
    uint32_t max_atoms;  // atoms
    uint32_t max_blob;   // blob bytes
    uint32_t max_seg;    // anchor segments (decode)
    uint32_t staged;     // 1: blob (and residue types) staged in smem; 0: large tier, the blob is written to global memory directly
    uint32_t stage_x;    // 1: the chain's coordinates are staged in smem by a bulk copy; 0: read from global memory (L1 / L2)
    uint32_t gws;        // 1: the workspace (atom offsets, angle arrays) lives in global memory too (k_encode_long)
    uint32_t threads;
    uint32_t smem;       // dynamic shared memory bytes
};

#define FCZ_NTIER 8
#define FCZ_MAX_CHUNKS 4096       // decode sub-batches per call
#define FCZ_SUB_RESIDUES 8000000u  // residues per decode sub-batch (bounds the workspace: ~80 B/residue)
// residue caps per tier; decode carries more per-residue state in shared memory (198 B vs 175 B), so its
// last staged tier is smaller.  Tier 6 keeps chain data in global memory, tier 7 (up to the format's 65535
// residues) its workspace as well.
// Caps are chosen at the occupancy steps of the per-residue shared-memory footprint (encode ~172 B/residue,
// 350-residue chains run 3 CTAs/SM).  Decode has its own four length tiers (dec_tier_of, below).
// Tier 3's cap is where EIGHT blocks of 128 threads still fit an SM's shared memory (8 x (26.5 KB + 1 KB) <= 228 KB) now
// that k_encode runs in 64 registers: measured against cap 384 / 72 registers / 7 blocks, k_encode 0.4131 -> 0.4025 ms on
// the 350-residue batch and 0.570 -> 0.559 ms per step on mixed lengths (profiles/r02_v15_ab.jsonl, r02_v16_ab_mixed.jsonl).
#ifndef FCZ_ENC_TIER3
#define FCZ_ENC_TIER3 352
#endif
static const uint32_t kEncTierRes[FCZ_NTIER] = {64, 128, 256, FCZ_ENC_TIER3, 640, 1280, 2720, 65535};

__host__ __device__ inline uint32_t align16(uint32_t x) { return (x + 15u) & ~15u; }

// ---- shared memory carve-up (offsets are computed identically on host and device)
struct EncSmem {
    uint32_t o_tab, o_misc, o_red, o_fl, o_type, o_aoff, o_ares, o_ang, o_x, o_b, total;
};
__host__ __device__ inline EncSmem enc_smem(const TierCfg& t) {
    EncSmem s;
    uint32_t o = 0;
    s.o_tab = o;  o += align16((uint32_t)offsetof(Tables, alt));  // natoms, name1, pred
    s.o_misc = o; o += 256;  // mbarrier, ticket, warp sums
    s.o_red = o;  o += align16(4u * FCZ_RED_FLOATS(24));  // blocks have at most 768 threads
    s.o_fl = o;   o += align16(4u * FCZ_FL_WORDS);
    s.o_type = o; o += t.staged ? align16(t.max_res) : 0u;
    s.o_aoff = o; o += t.gws ? 0u : align16(4u * (t.max_res + 1u));
    s.o_ares = o; o += t.gws ? 0u : align16(2u * t.max_atoms);  // side-chain atom -> residue; later the undecided-value list (enc_tier_list_cap)
    s.o_ang = o;  o += t.gws ? 0u : align16(24u * t.max_res);
    s.o_x = o;    o += t.stage_x ? align16(12u * t.max_atoms) + 32u : 0u;
    s.o_b = o;    o += t.staged ? align16(t.max_blob) + 32u : 0u;
    s.total = o;
    return s;
}
// entries of the undecided-value list + exact values (8 bytes each) that fit the side-chain map they alias
__host__ __device__ inline uint32_t enc_tier_list_cap(const TierCfg& t) {
    const uint32_t fit = align16(2u * t.max_atoms) / 8u, want = enc_list_cap(t.max_res);
    return want < fit ? want : fit;
}
static TierCfg make_tier(uint32_t max_res, bool staged, bool gws) {
    TierCfg t;
    t.max_res = max_res;
    t.staged = staged ? 1u : 0u;
    t.stage_x = t.staged;
    t.gws = gws ? 1u : 0u;
    t.max_atoms = 9u * max_res;
    t.max_blob = 17u * max_res + 1280u;
    t.max_seg = max_res / 10u + 8u;
    if (t.max_seg > 254u) t.max_seg = 254u;
    t.threads = (max_res <= 128u) ? 128u : 320u;
    if (!staged) {
        t.max_atoms = gws ? FCZ_MAX_ATOMS * max_res : 10u * max_res;
        t.max_blob = 0xFFFFFFFFu;
        t.max_seg = 254u;
        t.threads = 256u;
    }
    t.smem = 0;
    return t;
}
static void make_tiers(TierCfg* enc) {
    for (int i = 0; i < FCZ_NTIER; i++) {
        enc[i] = make_tier(kEncTierRes[i], i < FCZ_NTIER - 2, i == FCZ_NTIER - 1);
        enc[i].smem = enc_smem(enc[i]).total;
        enc[i].threads = 128u;  // settled at engine creation from the occupancy the footprint allows
    }
}

// ======================================================================================= device ctx

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 22)) __trap();  // a lost bulk copy must fail loudly, never hang the GPU
    }
}
// 1-D bulk async copy global -> shared (TMA engine), completion counted on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifdef FCZ_PHASE_TIMING
// Debug build only (libfcz_engine_timing.so): cycles spent between phase marks, summed over chains.
// ids: 0-3 encode phases, 4 encode stage/copy-out, 8-12 decode phases, 13 decode stage-in, 14 decode copy-out
__device__ unsigned long long g_phase_cycles[16];
__device__ unsigned long long g_phase_count[16];
#endif

struct DevCtx {
    int tid, nthr, lane, warp, nwarps;
    static constexpr int wsize = 32;
#ifdef FCZ_PHASE_TIMING
    long long t_last;
    __device__ __forceinline__ void mark(int id) {
        if (tid == 0) {
            long long t = clock64();
            atomicAdd(&g_phase_cycles[id], (unsigned long long)(t - t_last));
            atomicAdd(&g_phase_count[id], 1ull);
            t_last = t;
        }
    }
#else
    __device__ __forceinline__ void mark(int) {}
#endif
    uint32_t* wsum;   // [32] warp sums for the block scan
    uint64_t* bar;    // staging mbarrier
    uint32_t parity;  // its current phase
    bool staged;      // a bulk copy is in flight for this chain
    __device__ __forceinline__ void sync() { __syncthreads(); }
    __device__ __forceinline__ void wsync() { __syncwarp(); }
    __device__ __forceinline__ uint32_t atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
    __device__ __forceinline__ void stage_wait() {
        if (staged) mbar_wait(bar, parity);
    }
    __device__ __forceinline__ uint32_t excl_scan(uint32_t v) {
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t yv = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += yv;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += wsum[w];
        __syncthreads();
        return base + x - v;
    }
    __device__ __forceinline__ float wmin(float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    __device__ __forceinline__ float wmax(float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    // integer warp reductions (REDUX.MIN / REDUX.MAX) and shared / global atomics: floats go through fcz::ford
    __device__ __forceinline__ int32_t wmin_i(int32_t v) { return __reduce_min_sync(0xffffffffu, v); }
    __device__ __forceinline__ int32_t wmax_i(int32_t v) { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ void copy_out_same_phase(char* dst, const char* src, uint32_t bytes) const;  // = copy_out (below)
    __device__ __forceinline__ void atomic_min_u(uint32_t* p, uint32_t v) { atomicMin(p, v); }
    __device__ __forceinline__ void atomic_or_u(uint32_t* p, uint32_t v) { atomicOr(p, v); }
    __device__ __forceinline__ void atomic_min_i(int32_t* p, int32_t v) { atomicMin(p, v); }
    __device__ __forceinline__ void atomic_max_i(int32_t* p, int32_t v) { atomicMax(p, v); }
};

// Stage `bytes` bytes starting at global address `src` into shared memory so that the copy has
// the same 16-byte phase as the source: returns the shared address of byte 0.  The 16-byte
// aligned interior goes through ONE bulk async copy (issued by thread 0, completion on cx.bar);
// the <16-byte head and tail are fetched with plain loads.  buf must be 16-byte aligned and hold
// bytes + 32.
__device__ __forceinline__ uint8_t* stage_in(DevCtx& cx, uint8_t* buf, const uint8_t* src, uint32_t bytes) {
    const uint32_t mis = (uint32_t)((uintptr_t)src & 15u);
    uint8_t* dst0 = buf + mis;
    const uint32_t head = (16u - mis) & 15u;
    uint32_t body = 0;
    if (bytes > head) body = (bytes - head) & ~15u;
    if (cx.tid == 0) {
        // order this CTA's earlier generic-proxy accesses to the buffer before the async-proxy write
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (body) {
            mbar_expect_tx(cx.bar, body);
            bulk_g2s(dst0 + head, src + head, body, cx.bar);
        } else {
            mbar_arrive(cx.bar);
        }
    }
    const uint32_t hb = head < bytes ? head : bytes;
    for (uint32_t i = cx.tid; i < hb; i += cx.nthr) dst0[i] = src[i];
    for (uint32_t i = hb + body + cx.tid; i < bytes; i += cx.nthr) dst0[i] = src[i];
    cx.staged = true;
    return dst0;
}

// Copy `bytes` bytes from shared (same 16-byte phase as dst, see stage buffers) to global: the 16-byte aligned interior by
// ONE bulk async copy shared -> global (cp.async.bulk.global.shared::cta, SASS UBLKCP.G.S, issued by thread 0 after every
// thread has fenced its generic-proxy writes towards the async proxy), the <16-byte head and tail by plain stores.  Every
// thread of the block must call it; when it returns the image has been READ (thread 0 waited for the bulk group), so a
// barrier after it frees the buffer.  FCZ_BULK_STORE=0 builds the loop of 128-bit stores instead (A/B).
#ifndef FCZ_BULK_STORE
#define FCZ_BULK_STORE 1
#endif
__device__ __forceinline__ void copy_out(const DevCtx& cx, uint8_t* dst, const uint8_t* src, uint32_t bytes) {
    const uint32_t mis = (uint32_t)((uintptr_t)dst & 15u);
    const uint32_t head = (16u - mis) & 15u;
    const uint32_t hb = head < bytes ? head : bytes;
    uint32_t body = 0;
    if (bytes > head) body = (bytes - head) & ~15u;
#if FCZ_BULK_STORE
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (cx.tid == 0 && body) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head), "r"(smem_u32(src + head)), "r"(body) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    for (uint32_t i = cx.tid; i < hb; i += cx.nthr) dst[i] = src[i];
    for (uint32_t i = hb + body + cx.tid; i < bytes; i += cx.nthr) dst[i] = src[i];
    if (cx.tid == 0 && body) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#else
    for (uint32_t i = cx.tid; i < hb; i += cx.nthr) dst[i] = src[i];
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (uint32_t i = cx.tid; i < (body >> 4); i += cx.nthr) d4[i] = s4[i];
    for (uint32_t i = hb + body + cx.tid; i < bytes; i += cx.nthr) dst[i] = src[i];
#endif
}

__device__ __forceinline__ void DevCtx::copy_out_same_phase(char* dst, const char* src, uint32_t bytes) const {
    copy_out(*this, reinterpret_cast<uint8_t*>(dst), reinterpret_cast<const uint8_t*>(src), bytes);
}

// ========================================================================================== encode

// What k_encode needs to know about a chain before it touches its data, written by k_enc_plan at the chain's position in
// its tier list (one 48-byte load instead of list -> three offset arrays at the top of every trip).
struct alignas(16) EncDesc {
    uint32_t c, r0, L, A;
    uint32_t t0, T, pad0, pad1;
    uint64_t a0, pad2;
};
static_assert(sizeof(EncDesc) == 48, "EncDesc is three 16-byte words");

struct EncArgs {
    const uint32_t* res_off;
    const uint64_t* atom_off;
    const uint32_t* title_off;
    const uint8_t* res_type;
    const float* bfactor;
    const float* xyz;
    const char* titles;
    const fcz_chain_meta* meta;
    const uint64_t* blob_off;
    uint8_t* bytes;
    const struct EncDesc* desc;  // per position of `list` (device-planned batches), or null: the kernel reads the offset arrays
    const uint32_t* list;   // chains of this tier
    const uint32_t* count;  // their number (device-planned batches) ...
    uint32_t count_val;     // ... or by value when count == nullptr (host-planned batches)
    uint32_t* ticket;
    const Tables* tables;
    int32_t b;
    TierCfg cfg;
    uint32_t term;        // 1: one NUL byte follows every blob (fcz_opts.terminate_blobs)
    uint32_t prefetch_next;  // 1: pull the block's next chain towards L2 while the current one is encoded
    uint8_t* gws;         // k_encode_long: per-block workspace in global memory, gws_stride bytes each
    uint64_t gws_stride;
    uint32_t gws_max_res; // residues the workspace was sized for
};

// Plan: one warp per chain.  Validates the chain, computes its blob size and tier.
struct PlanOut {
    uint32_t* v0;          // per-chain value 0 (encode: blob bytes; decode: residues)
    uint32_t* v1;          // decode: atoms
    uint32_t* v2;          // decode: title bytes
    uint32_t* v3;          // decode: anchors (segment scratch slots)
    int32_t* status;       // per-chain status (engine copy)
    uint32_t* tier_count;  // [FCZ_NTIER]
    uint32_t* tier_list;   // [FCZ_NTIER][n]
    EncDesc* enc_desc;     // [FCZ_NTIER][n] or null (encode plan only)
};

struct TierTable {
    TierCfg t[FCZ_NTIER];
};

__device__ __forceinline__ int pick_tier(const TierTable& tt, uint32_t L, uint32_t A, uint32_t blob, uint32_t nseg) {
    for (int i = 0; i < FCZ_NTIER; i++) {
        const TierCfg& t = tt.t[i];
        if (L <= t.max_res && A <= t.max_atoms && blob <= t.max_blob && nseg <= t.max_seg) return i;
    }
    return -1;
}

__global__ void k_enc_plan(uint32_t n, const uint32_t* res_off, const uint64_t* atom_off, const uint32_t* title_off,
                           const uint8_t* res_type, int32_t b, uint32_t term, const Tables* tb, TierTable tt, PlanOut po) {
    const uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    const uint32_t r0 = res_off[c], L = res_off[c + 1] - r0;
    const uint64_t A = atom_off[c + 1] - atom_off[c];
    const uint32_t T = title_off[c + 1] - title_off[c];
    int status = FCZ_OK;
    uint32_t sum = 0, bad = 0;
    for (uint32_t r = lane; r < L; r += 32) {
        uint32_t code = res_type[r0 + r];
        uint32_t na = code < FCZ_NUM_CODES ? natoms_packed(code) : 0u;
        bad |= (na == 0u);
        sum += na;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    uint32_t size = 0;
    int tier = -1;
    if (L < 2u || L > 65535u || b < 1) status = FCZ_E_LIMIT;
    else if (bad) status = FCZ_E_RESIDUE;
    else if ((uint64_t)sum != A) status = FCZ_E_ARG;
    else {
        const int na = anchor_count(L, b);
        if (na > 255) status = FCZ_E_LIMIT;
        else {
            Layout y = make_layout(L, sum - 3u * L, T, (uint32_t)na);
            size = y.size + term;
            tier = pick_tier(tt, L, sum, size, (uint32_t)na - 1u);
            if (tier < 0) { status = FCZ_E_LIMIT; size = 0; }
        }
    }
    if (lane == 0) {
        po.v0[c] = size;
        po.status[c] = status;
        if (tier >= 0) {
            uint32_t pos = atomicAdd(&po.tier_count[tier], 1u);
            po.tier_list[(size_t)tier * n + pos] = c;
            if (po.enc_desc) {
                EncDesc q;
                q.c = c; q.r0 = r0; q.L = L; q.A = sum; q.t0 = title_off[c]; q.T = T; q.pad0 = 0u; q.pad1 = 0u; q.a0 = atom_off[c]; q.pad2 = 0ull;
                po.enc_desc[(size_t)tier * n + pos] = q;
            }
            if (tt.t[tier].gws) atomicMax(&po.tier_count[2 * FCZ_NTIER], L);  // sizes the long tier's global workspace
        }
    }
}

#ifndef FCZ_ENC_MAXREG
#define FCZ_ENC_MAXREG 64  // 1024 threads per SM: eight blocks of 128 (96 bytes of spill stores; 72 registers = seven blocks, measured slower)
#endif
__global__ void __maxnreg__(FCZ_ENC_MAXREG) k_encode(EncArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const EncSmem so = enc_smem(a.cfg);
    Tables* tb = reinterpret_cast<Tables*>(smem + so.o_tab);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + so.o_misc);
    uint32_t* s_ticket = reinterpret_cast<uint32_t*>(smem + so.o_misc + 8);
    DevCtx cx;
    cx.tid = threadIdx.x; cx.nthr = blockDim.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.nwarps = blockDim.x >> 5;
    cx.wsum = reinterpret_cast<uint32_t*>(smem + so.o_misc + 64);
    cx.bar = bar; cx.parity = 0; cx.staged = false;
    // tables prefix (natoms, alt, pred) -> shared
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tb);
        for (uint32_t i = cx.tid; i < (uint32_t)offsetof(Tables, alt) / 4u; i += cx.nthr) dst[i] = src[i];
    }
    if (cx.tid == 0) mbar_init(bar, 1);
    __syncthreads();
    const uint32_t count = a.count ? *a.count : a.count_val;
    // Tickets run one chain ahead: while chain t is encoded, thread 0 already holds the ticket of the block's next chain
    // (slots alternate, so a slot is rewritten two barriers after its last reader), and the next chain's coordinates
    // are pulled towards L2 -- the bulk copy (or the first loads) of the next trip then find them there.
    if (cx.tid == 0) s_ticket[0] = atomicAdd(a.ticket, 1u);
    __syncthreads();
    uint32_t t = s_ticket[0];
    for (uint32_t trip = 0; t < count; trip++) {
        if (cx.tid == 0) s_ticket[(trip + 1u) & 1u] = atomicAdd(a.ticket, 1u);
#ifdef FCZ_PHASE_TIMING
        cx.t_last = clock64();
#endif
        uint32_t c, r0, L, A, t0, T;
        uint64_t a0;
        if (a.desc) {
            const EncDesc q = a.desc[t];
            c = q.c; r0 = q.r0; L = q.L; A = q.A; t0 = q.t0; T = q.T; a0 = q.a0;
        } else {
            c = a.list[t];
            r0 = a.res_off[c]; L = a.res_off[c + 1] - r0;
            a0 = a.atom_off[c];
            A = (uint32_t)(a.atom_off[c + 1] - a0);
            t0 = a.title_off[c]; T = a.title_off[c + 1] - t0;
        }
        const uint64_t b0 = a.blob_off[c];
        const uint32_t size = (uint32_t)(a.blob_off[c + 1] - b0);

        EncChain ch;
        ch.L = L; ch.A = A; ch.title_len = T; ch.b = a.b;
        ch.bfac = a.bfactor + r0;
        ch.title = a.titles + t0;
        ch.meta = a.meta + c;
        ch.aoff = reinterpret_cast<uint32_t*>(smem + so.o_aoff);
        ch.sres = reinterpret_cast<uint16_t*>(smem + so.o_ares);
        ch.ang = reinterpret_cast<float*>(smem + so.o_ang);
        ch.red = reinterpret_cast<float*>(smem + so.o_red);
        ch.fl = reinterpret_cast<uint32_t*>(smem + so.o_fl);
        ch.tbg = a.tables;
        ch.list_cap = enc_tier_list_cap(a.cfg);
        ch.list = reinterpret_cast<uint32_t*>(smem + so.o_ares);
        ch.xe = reinterpret_cast<float*>(smem + so.o_ares) + ch.list_cap;
        uint8_t* gdst = a.bytes + b0;
        uint8_t* sB = nullptr;
        if (a.cfg.staged) {
            if (a.cfg.stage_x) {
                const uint8_t* gx = reinterpret_cast<const uint8_t*>(a.xyz + 3u * a0);
                ch.X = reinterpret_cast<const float*>(stage_in(cx, smem + so.o_x, gx, 12u * A));
            } else {
                // coordinates stay in global memory (L1 / L2): the block's first chain is pulled towards L2 here, every
                // later one was prefetched a trip ahead (below)
                ch.X = a.xyz + 3u * a0;
                if (trip == 0u || !a.prefetch_next) {
                    const char* gx = reinterpret_cast<const char*>(ch.X);
                    for (uint32_t off = cx.tid * 128u; off < 12u * A; off += cx.nthr * 128u) prefetch_l2(gx + off);
                }
                cx.staged = false;
            }
            uint8_t* st = smem + so.o_type;
            for (uint32_t i = cx.tid; i < L; i += cx.nthr) st[i] = a.res_type[r0 + i];
            ch.type = st;
            sB = smem + so.o_b + ((uintptr_t)gdst & 15u);
            ch.B = sB;
            __syncthreads();  // staged types (and the next ticket) visible before phase 1
            {
                const uint32_t tn = s_ticket[(trip + 1u) & 1u];
                if (a.prefetch_next && tn < count) {
                    const uint32_t cn = a.list[tn];
                    const uint64_t an = a.atom_off[cn];
                    const uint32_t bytes = 12u * (uint32_t)(a.atom_off[cn + 1] - an);
                    const char* gx = reinterpret_cast<const char*>(a.xyz + 3u * an);
                    for (uint32_t off = cx.tid * 128u; off < bytes; off += cx.nthr * 128u) prefetch_l2(gx + off);
                    if (cx.tid < 32u) {
                        const uint32_t rn = a.res_off[cn], Ln = a.res_off[cn + 1] - rn;
                        for (uint32_t off = cx.tid * 128u; off < Ln; off += 32u * 128u) prefetch_l2(a.res_type + rn + off);
                        for (uint32_t off = cx.tid * 128u; off < 4u * Ln; off += 32u * 128u) prefetch_l2(reinterpret_cast<const char*>(a.bfactor + rn) + off);
                    }
                }
            }
        } else {
            ch.X = a.xyz + 3u * a0;
            ch.type = a.res_type + r0;
            ch.B = gdst;
            cx.staged = false;
        }
        __builtin_assume(__isShared(tb));
        __builtin_assume(__isShared(ch.aoff));
        __builtin_assume(__isShared(ch.sres));
        __builtin_assume(__isShared(ch.ang));
        __builtin_assume(__isShared(ch.red));
        __builtin_assume(__isShared(ch.fl));
        __builtin_assume(__isShared(ch.list));
        __builtin_assume(__isShared(ch.xe));
        // Three instances of the codec, by the address space of the coordinates and of the blob.  The coordinate pointer
        // of the global-memory instance is re-derived from the kernel parameter instead of being given an assumption (two
        // different address-space assumptions on one SSA value leaked across the branches: LDG on a shared address).
        if (a.cfg.staged && a.cfg.stage_x) {
            __builtin_assume(__isShared(ch.X));
            __builtin_assume(__isShared(ch.type));
            __builtin_assume(__isShared(ch.B));
            encode_chain(cx, tb, ch);
        } else if (a.cfg.staged) {
#ifndef FCZ_ANALYSE_STAGED_ONLY
            ch.X = a.xyz + 3u * a0;  // kernel parameter: known to be global memory
            __builtin_assume(__isShared(ch.type));
            __builtin_assume(__isShared(ch.B));
            encode_chain(cx, tb, ch);
#endif
        } else {
#ifndef FCZ_ANALYSE_STAGED_ONLY  // (tools/sass_static.py: count one copy of the codec)
            encode_chain(cx, tb, ch);
#endif
        }
        if (a.term && cx.tid == 0) (a.cfg.staged ? sB : gdst)[size - 1u] = 0;  // encode_chain ended with a barrier
        if (a.cfg.staged) {
            copy_out(cx, gdst, sB, size);  // (synchronises the block first: the terminator is in)
            cx.mark(4);
            if (a.cfg.stage_x) cx.parity ^= 1u;
            cx.staged = false;
        }
        __syncthreads();  // shared buffers free for the next chain
        t = s_ticket[(trip + 1u) & 1u];
    }
}

// The long tier (2721 .. 65535 residues): chain data AND workspace stay in global memory (per-block slices of
// a.gws, sized for the tier's longest chain); otherwise the same persistent ticket loop as k_encode.
__host__ __device__ inline uint64_t enc_gws_bytes(uint32_t max_res) {
    return (uint64_t)align16(4u * (max_res + 1u)) + align16(2u * (FCZ_MAX_ATOMS - 3u) * max_res) + align16(24u * max_res) +
           align16(8u * enc_list_cap(max_res));
}
__global__ void __launch_bounds__(1024) k_encode_long(EncArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const EncSmem so = enc_smem(a.cfg);
    Tables* tb = reinterpret_cast<Tables*>(smem + so.o_tab);
    uint32_t* s_ticket = reinterpret_cast<uint32_t*>(smem + so.o_misc + 8);
    DevCtx cx;
    cx.tid = threadIdx.x; cx.nthr = blockDim.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.nwarps = blockDim.x >> 5;
    cx.wsum = reinterpret_cast<uint32_t*>(smem + so.o_misc + 64);
    cx.bar = nullptr; cx.parity = 0; cx.staged = false;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tb);
        for (uint32_t i = cx.tid; i < (uint32_t)offsetof(Tables, alt) / 4u; i += cx.nthr) dst[i] = src[i];
    }
    __syncthreads();
    uint8_t* ws = a.gws + (uint64_t)blockIdx.x * a.gws_stride;
    const uint32_t count = a.count ? *a.count : a.count_val;
    if (blockDim.x > 768u) __trap();  // FCZ_RED_FLOATS(24)
    for (;;) {
        if (cx.tid == 0) *s_ticket = atomicAdd(a.ticket, 1u);
        __syncthreads();
        const uint32_t t = *s_ticket;
        if (t >= count) break;
        const uint32_t c = a.list[t];
        const uint32_t r0 = a.res_off[c], L = a.res_off[c + 1] - r0;
        const uint64_t a0 = a.atom_off[c];
        const uint32_t t0 = a.title_off[c];
        if (L > a.gws_max_res) __trap();  // the plan sized the workspace from these very chains
        EncChain ch;
        ch.L = L; ch.A = (uint32_t)(a.atom_off[c + 1] - a0); ch.title_len = a.title_off[c + 1] - t0; ch.b = a.b;
        ch.bfac = a.bfactor + r0;
        ch.title = a.titles + t0;
        ch.meta = a.meta + c;
        ch.aoff = reinterpret_cast<uint32_t*>(ws);
        ch.sres = reinterpret_cast<uint16_t*>(ws + align16(4u * (a.gws_max_res + 1u)));
        ch.ang = reinterpret_cast<float*>(ws + align16(4u * (a.gws_max_res + 1u)) + align16(2u * (FCZ_MAX_ATOMS - 3u) * a.gws_max_res));
        ch.red = reinterpret_cast<float*>(smem + so.o_red);
        ch.fl = reinterpret_cast<uint32_t*>(smem + so.o_fl);
        ch.tbg = a.tables;
        ch.list_cap = enc_list_cap(a.gws_max_res);
        ch.list = reinterpret_cast<uint32_t*>(ws + align16(4u * (a.gws_max_res + 1u)) + align16(2u * (FCZ_MAX_ATOMS - 3u) * a.gws_max_res) + align16(24u * a.gws_max_res));
        ch.xe = reinterpret_cast<float*>(ch.list) + ch.list_cap;
        ch.X = a.xyz + 3u * a0;
        ch.type = a.res_type + r0;
        ch.B = a.bytes + a.blob_off[c];
        encode_chain(cx, tb, ch);
        if (a.term && cx.tid == 0) ch.B[(uint32_t)(a.blob_off[c + 1] - a.blob_off[c]) - 1u] = 0;
        __syncthreads();
    }
}

// ========================================================================================== decode

// validate_only: host-planned batches -- chains [c0, n) are checked, only po.status is written (and only
// for chains the host has not already rejected).
__global__ void k_dec_plan(uint32_t c0, uint32_t n, const uint64_t* blob_off, const uint8_t* bytes, const Tables* tb, TierTable tt,
                           PlanOut po, int validate_only) {
    const uint32_t c = c0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n) return;
    if (validate_only && po.status[c] != FCZ_OK) return;
    const uint8_t* blob = bytes + blob_off[c];
    const uint64_t len = blob_off[c + 1] - blob_off[c];
    int status = FCZ_OK;
    uint32_t L = 0, A = 0, T = 0;
    if (len < HDR_BYTES || blob[0] != 'F' || blob[1] != 'C' || blob[2] != 'M' || blob[3] != 'P') {
        status = FCZ_E_MAGIC;
    } else {
        L = get_u16(blob + OFF_NRES);
        T = get_u32(blob + OFF_LENTITLE);
        const uint32_t nsc = get_u32(blob + OFF_NSC), na = blob[OFF_NANCHOR];
        Layout y = make_layout(L, nsc, T, na);
        if (L < 2u || na < 2u || (uint64_t)T > len || (uint64_t)nsc > len || (uint64_t)y.size > len) {
            status = FCZ_E_TRUNCATED;
        } else {
            uint32_t sum = 0, bad = 0;
            for (uint32_t r = lane; r < L; r += 32) {
                uint32_t nat = natoms_packed(blob[y.o_rec + 8u * r] >> 3);
                bad |= (nat == 0u);
                sum += nat;
            }
            // anchors: 0 = first, non-decreasing, last = L-1
            for (uint32_t i = lane; i < na; i += 32) {
                uint32_t ai = get_u32(blob + y.o_aidx + 4u * i);
                uint32_t prev = i ? get_u32(blob + y.o_aidx + 4u * (i - 1u)) : 0u;
                if (ai < prev || ai >= L || (i == 0u && ai != 0u) || (i == na - 1u && ai != L - 1u)) bad |= 2u;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                bad |= __shfl_xor_sync(0xffffffffu, bad, o);
            }
            if (bad & 1u) status = FCZ_E_RESIDUE;
            else if ((bad & 2u) || sum - 3u * L != nsc) status = FCZ_E_TRUNCATED;
            else {
                A = sum;
            }
        }
    }
    if (status != FCZ_OK) { L = 0; A = 0; T = 0; }
    if (validate_only) {
        if (lane == 0 && status != FCZ_OK) po.status[c] = status;
        return;
    }
    if (lane == 0) {
        po.v0[c] = L; po.v1[c] = A; po.v2[c] = T;
        po.v3[c] = (status == FCZ_OK) ? (uint32_t)blob[OFF_NANCHOR] : 0u;
        po.status[c] = status;
    }
}

// ================================================================================== decode kernels
// Decode runs phase by phase over all chains of a length tier (fcz_codec.h: dec_unpack, dec_passes, dec_stitch_core,
// dec_blend, dec_side), because a CTA that walks one chain through its serial phase (the stitch: one thread) would
// leave its SM idle.  Two forms: the shared-memory pipeline k_dec_front -> k_dec_stitch_t -> k_dec_back (working set
// of a chain in shared memory, hand-over through global scratch), and, for chains whose working set does not fit
// 227 KB, one kernel per phase with the whole workspace in global memory (k_dec_unpack ... k_dec_side).

// Everything the decode kernels need to know about a chain before they touch its data, gathered by k_plan_chunks at the
// chain's position in the tier list: one 64-byte load where the kernels would chase list -> six offset arrays (two
// dependent round trips at the start of every block).
struct alignas(16) DecDesc {
    uint32_t c, r0, L, A;
    uint32_t t0, T, g0, nA;
    uint64_t a0, b0;
    uint32_t blob_len;
    int32_t status;
    uint32_t pad[2];
};
static_assert(sizeof(DecDesc) == 64, "DecDesc is four 16-byte words");

struct Dec2Args {
    const DecDesc* desc;      // per position of `list` (device-planned batches), or null: the kernels read the offset arrays
    const uint64_t* blob_off;
    const uint8_t* bytes;
    const uint32_t* res_off;
    const uint64_t* atom_off;
    const uint32_t* title_off;
    const uint32_t* seg_off;  // [n+1] prefix of anchors per chain = first segment-scratch slot
    uint8_t* res_type;
    float* bfactor;
    float* xyz;
    char* titles;
    fcz_chain_meta* meta;
    const int32_t* status;
    const Tables* tables;
    int32_t use_alt;
    uint32_t c0, c1;          // chains of this sub-batch
    const uint32_t* list;     // chains of this launch: a length tier of the sub-batch (block i = list[i])
    uint32_t count;
    uint32_t r_base, s_base;  // res_off[c0], seg_off[c0]
    uint32_t* aoff;           // workspace, indexed from the sub-batch's first residue / segment slot
    uint8_t* segid;
    cs* tor;
    cs* ang;
    float* rev;
    float* seg;
    // shared-memory path (k_dec_front / k_dec_stitch_soa / k_dec_back)
    float* loc;               // [9 * residues] forward-pass backbone atoms in segment-local coordinates
    uint32_t stitch_group;    // chains per block of k_dec_stitch_t
    uint32_t max_L, max_anchor, max_blob, max_atoms;  // of the sub-batch: shared-memory carve-up
};

__device__ __forceinline__ bool dec2_chain(const Dec2Args& a, uint32_t c, DecChain& ch) {
    if (a.status[c] != FCZ_OK) return false;
    const uint8_t* blob = a.bytes + a.blob_off[c];
    ch.blob = blob;
    ch.y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
    ch.use_alt = a.use_alt;
    const uint32_t r0 = a.res_off[c], rr = r0 - a.r_base;
    ch.out_xyz = a.xyz + 3u * a.atom_off[c];
    ch.out_type = a.res_type + r0;
    ch.out_bfac = a.bfactor + r0;
    ch.out_meta = a.meta + c;
    ch.out_title = a.titles ? a.titles + a.title_off[c] : nullptr;
    ch.aoff = a.aoff + rr + (c - a.c0);
    ch.segid = a.segid + rr;
    ch.tor = a.tor + 3u * (size_t)rr;
    ch.ang = a.ang + 3u * (size_t)rr;
    ch.rev = a.rev + 9u * (size_t)rr;
    ch.loc = nullptr;
    ch.order = nullptr; ch.bins = nullptr; ch.codes = nullptr; ch.sc = nullptr;
    ch.seg = a.seg + (size_t)(a.seg_off[c] - a.s_base) * FCZ_SEG_FLOATS;
    return true;
}

struct ThreadCtx {  // one thread = one chain (stitch)
    static constexpr int tid = 0, nthr = 1, lane = 0, warp = 0, nwarps = 1, wsize = 1;
    __device__ __forceinline__ void sync() {}
    __device__ __forceinline__ void wsync() {}
};

__device__ __forceinline__ DevCtx block_ctx(uint32_t* wsum) {
    DevCtx cx;
    cx.tid = threadIdx.x; cx.nthr = blockDim.x; cx.lane = threadIdx.x & 31; cx.warp = threadIdx.x >> 5;
    cx.nwarps = blockDim.x >> 5;
    cx.wsum = wsum; cx.bar = nullptr; cx.parity = 0; cx.staged = false;
    return cx;
}

__global__ void __launch_bounds__(256) k_dec_unpack(Dec2Args a) {  // block per chain
    __shared__ uint32_t wsum[32];
    DecChain ch;
    if (!dec2_chain(a, a.list[blockIdx.x], ch)) return;
    DevCtx cx = block_ctx(wsum);
    dec_unpack(cx, a.tables, ch);
}
__global__ void __launch_bounds__(128) k_dec_passes(Dec2Args a) {  // block per chain: forward items in two warps, reverse items in two
    DecChain ch;
    if (!dec2_chain(a, a.list[blockIdx.x], ch)) return;
    DevCtx cx = block_ctx(nullptr);
    dec_passes(cx, a.tables, ch);
}
__global__ void __launch_bounds__(128) k_dec_blend(Dec2Args a) {  // block per chain, thread per backbone atom
    DecChain ch;
    if (!dec2_chain(a, a.list[blockIdx.x], ch)) return;
    DevCtx cx = block_ctx(nullptr);
    dec_blend(cx, a.tables, ch);
}
__global__ void __launch_bounds__(192) k_dec_side(Dec2Args a) {  // block per chain, thread per residue pair
    DecChain ch;
    if (!dec2_chain(a, a.list[blockIdx.x], ch)) return;
    DevCtx cx = block_ctx(nullptr);
    dec_side(cx, a.tables, ch);
}

// ---------------------------------------------------------------- shared-memory decode, three kernels
// front (block per chain): blob staged by one bulk copy; unpack + both NeRF passes with records, (cos,sin)
//   tables and segment scratch in shared memory; hands over the local/reverse backbone atoms and the atom
//   offsets in global memory and the segment scratch as a structure-of-arrays over the sub-batch's chains.
// stitch (thread per chain): the serial walk over segments, coalesced over chains, paid once per sub-batch.
// back (block per chain): blend + side chains into a shared-memory image of the chain's coordinates, then
//   one 128-bit copy-out.
struct FrontSmem { uint32_t o_seg, o_tor, o_ang, o_aoff, o_segid, o_blob, total; };
__host__ __device__ inline uint32_t up16(uint32_t v) { return (v + 15u) & ~15u; }
__host__ __device__ inline FrontSmem front_smem(uint32_t max_L, uint32_t max_anchor, uint32_t max_blob) {
    FrontSmem o;
    uint32_t p = 256;  // mbarrier + warp sums
    o.o_seg = p; p += up16(4u * FCZ_SEG_FLOATS * (max_anchor + 1u));
    o.o_tor = p; p += up16(24u * max_L);
    o.o_ang = p; p += up16(24u * max_L);
    o.o_aoff = p; p += up16(4u * (max_L + 1u));
    o.o_segid = p; p += up16(max_L);
    o.o_blob = p; p += up16(max_blob) + 32u;
    o.total = p;
    return o;
}
struct BackSmem { uint32_t o_seg, o_aoff, o_segid, o_codes, o_order, o_out, total; };
__host__ __device__ inline BackSmem back_smem(uint32_t max_L, uint32_t max_anchor, uint32_t max_atoms) {
    BackSmem o;
    uint32_t p = 0;
    o.o_seg = p; p += up16(4u * FCZ_SEG_FLOATS * (max_anchor + 1u));
    o.o_aoff = p; p += up16(4u * (max_L + 1u));
    o.o_segid = p; p += up16(max_L);
    o.o_codes = p; p += up16(max_L);
    o.o_order = p; p += up16(2u * max_L) + 128u;  // sorted residues + 32 bins
    o.o_out = p; p += up16(12u * max_atoms) + 32u;
    o.total = p;
    return o;
}

__global__ void __launch_bounds__(1024) k_dec_front(Dec2Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    // everything about the chain: from its descriptor (one load), else from the offset arrays (list, then independent
    // loads: two round trips); the plan has already checked them against the blob header
    uint32_t c, r0, L, A, t0, T, g0, nA;
    uint64_t b0, b1;
    int32_t st;
    if (a.desc) {
        const DecDesc q = a.desc[blockIdx.x];
        c = q.c; r0 = q.r0; L = q.L; A = q.A; t0 = q.t0; T = q.T; g0 = q.g0; nA = q.nA; b0 = q.b0; b1 = q.b0 + q.blob_len; st = q.status;
    } else {
        c = a.list[blockIdx.x];
        st = a.status[c];
        b0 = a.blob_off[c]; b1 = a.blob_off[c + 1];
        r0 = a.res_off[c]; L = a.res_off[c + 1] - r0;
        A = (uint32_t)(a.atom_off[c + 1] - a.atom_off[c]);
        t0 = a.title_off[c]; T = a.title_off[c + 1] - t0;
        g0 = a.seg_off[c]; nA = a.seg_off[c + 1] - g0;
    }
    if (st != FCZ_OK) return;
    const FrontSmem so = front_smem(a.max_L, a.max_anchor, a.max_blob);
    DevCtx cx = block_ctx(reinterpret_cast<uint32_t*>(smem + 64));
    cx.bar = reinterpret_cast<uint64_t*>(smem);
#ifdef FCZ_PHASE_TIMING
    cx.t_last = clock64();
#endif
    if (cx.tid == 0) mbar_init(cx.bar, 1);
    __syncthreads();
    const uint8_t* gblob = a.bytes + b0;
    const uint32_t len = (uint32_t)(b1 - b0);
    DecChain ch;
    ch.y = make_layout(L, A - 3u * L, T, nA);
    const uint32_t size = ch.y.size < len ? ch.y.size : len;
    ch.blob = stage_in(cx, smem + so.o_blob, gblob, size);
    ch.use_alt = a.use_alt;
    const uint32_t rr = r0 - a.r_base;
    ch.out_type = a.res_type + r0;
    ch.out_bfac = a.bfactor + r0;
    ch.out_meta = a.meta + c;
    ch.out_title = a.titles ? a.titles + t0 : nullptr;
    ch.aoff = reinterpret_cast<uint32_t*>(smem + so.o_aoff);
    ch.segid = smem + so.o_segid;
    ch.tor = reinterpret_cast<cs*>(smem + so.o_tor);
    ch.ang = reinterpret_cast<cs*>(smem + so.o_ang);
    ch.seg = reinterpret_cast<float*>(smem + so.o_seg);
    ch.rev = a.rev + 9u * (size_t)rr;
    ch.out_xyz = a.loc + 9u * (size_t)rr;  // forward atoms of residue r at 9r: see the offsets below
    ch.loc = nullptr;
    ch.order = nullptr; ch.bins = nullptr; ch.codes = nullptr; ch.sc = nullptr;
    __builtin_assume(__isShared(ch.blob));
    __builtin_assume(__isShared(ch.aoff));
    __builtin_assume(__isShared(ch.segid));
    __builtin_assume(__isShared(ch.tor));
    __builtin_assume(__isShared(ch.ang));
    __builtin_assume(__isShared(ch.seg));
    cx.stage_wait();
    cx.mark(13);
    dec_unpack(cx, a.tables, ch);
    __syncthreads();
    uint32_t* g_aoff = a.aoff + rr + (c - a.c0);
    for (uint32_t r = cx.tid; r <= L; r += cx.nthr) g_aoff[r] = ch.aoff[r];
    __syncthreads();
    for (uint32_t r = cx.tid; r <= L; r += cx.nthr) ch.aoff[r] = 3u * r;  // compact backbone-only slots for the passes
    __syncthreads();
    cx.mark(8);
    dec_passes(cx, a.tables, ch);
    cx.mark(9);
    // (copy_out fences every thread's writes and synchronises the block before the copy is issued)
    // the segment scratch (352 bytes per segment, 16-byte aligned on both sides) leaves by one bulk async copy
    copy_out(cx, reinterpret_cast<uint8_t*>(a.seg + (size_t)(g0 - a.s_base) * FCZ_SEG_FLOATS), reinterpret_cast<const uint8_t*>(ch.seg),
             nA * FCZ_SEG_FLOATS * 4u);
    cx.mark(14);
}

// stitch: block = stitch_group chains, one thread per chain walks the segments.  The fields it reads are first
// gathered (all threads, coalesced over each chain's scratch) into shared memory, packed 46 floats per slot with
// an odd per-chain stride (conflict-free for thread-per-chain access); S and T come back the same way.
__global__ void __launch_bounds__(1024) k_dec_stitch_t(Dec2Args a) {
    extern __shared__ __align__(16) float sm[];
    const uint32_t G = a.stitch_group, cstride = (a.max_anchor * SegPacked::N) | 1u;
    const uint32_t ib = blockIdx.x * G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (uint32_t j = warp; j < G; j += nwarps) {
        if (ib + j >= a.count) continue;
        uint32_t s0, nA;
        if (a.desc) { s0 = a.desc[ib + j].g0; nA = a.desc[ib + j].nA; }
        else { const uint32_t c = a.list[ib + j]; s0 = a.seg_off[c]; nA = a.seg_off[c + 1] - s0; }
        const float* src = a.seg + (size_t)(s0 - a.s_base) * FCZ_SEG_FLOATS;
        float* dst = sm + (size_t)j * cstride;
        const uint32_t ne = nA * SegPacked::N;
        for (uint32_t base = lane; base < ne; base += 32u * 8u) {  // eight loads in flight per lane
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t e = base + 32u * u;
                const uint32_t s = e / SegPacked::N, k = e - s * SegPacked::N;
                // packed order: I[3] F[12] TAIL[9] A[9] CS[13]
                const uint32_t f = k < 3u ? SEG_I + k : (k < 15u ? SEG_F + (k - 3u) : (k < 24u ? SEG_TAIL + (k - 15u) : (k < 33u ? SEG_A + (k - 24u) : SEG_CS + (k - 33u))));
                v[u] = e < ne ? __ldg(src + s * FCZ_SEG_FLOATS + f) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t e = base + 32u * u;
                if (e < ne) dst[e] = v[u];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < G && ib + threadIdx.x < a.count) {
        uint32_t nA;
        if (a.desc) nA = a.desc[ib + threadIdx.x].nA;
        else { const uint32_t c = a.list[ib + threadIdx.x]; nA = a.seg_off[c + 1] - a.seg_off[c]; }
        dec_stitch_core<SegPacked>(sm + (size_t)threadIdx.x * cstride, 1, (int)nA - 1);
    }
    __syncthreads();
    for (uint32_t j = warp; j < G; j += nwarps) {
        if (ib + j >= a.count) continue;
        uint32_t s0, nA;
        if (a.desc) { s0 = a.desc[ib + j].g0; nA = a.desc[ib + j].nA; }
        else { const uint32_t c = a.list[ib + j]; s0 = a.seg_off[c]; nA = a.seg_off[c + 1] - s0; }
        float* out = a.seg + (size_t)(s0 - a.s_base) * FCZ_SEG_FLOATS;
        const float* srcp = sm + (size_t)j * cstride;
        for (uint32_t e = lane; e < nA * 21u; e += 32) {  // S[9] T[12] are adjacent in the full layout
            const uint32_t s = e / 21u, k = e - s * 21u;
            out[s * FCZ_SEG_FLOATS + k] = srcp[s * SegPacked::N + (k < 9u ? SegPacked::S + k : SegPacked::T + (k - 9u))];
        }
    }
}

__global__ void __launch_bounds__(1024) k_dec_back(Dec2Args a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t wsum[32];
    uint32_t c, r0, L, A, T, g0, nA;
    uint64_t b0, a0;
    int32_t st;
    if (a.desc) {
        const DecDesc q = a.desc[blockIdx.x];
        c = q.c; r0 = q.r0; L = q.L; A = q.A; T = q.T; g0 = q.g0; nA = q.nA; b0 = q.b0; a0 = q.a0; st = q.status;
    } else {
        c = a.list[blockIdx.x];
        st = a.status[c];
        b0 = a.blob_off[c];
        r0 = a.res_off[c]; L = a.res_off[c + 1] - r0;
        a0 = a.atom_off[c];
        A = (uint32_t)(a.atom_off[c + 1] - a0);
        T = a.title_off[c + 1] - a.title_off[c];
        g0 = a.seg_off[c]; nA = a.seg_off[c + 1] - g0;
    }
    if (st != FCZ_OK) return;
    const BackSmem so = back_smem(a.max_L, a.max_anchor, a.max_atoms);
    DevCtx cx = block_ctx(wsum);
#ifdef FCZ_PHASE_TIMING
    cx.t_last = clock64();
#endif
    const uint8_t* blob = a.bytes + b0;
    DecChain ch;
    ch.blob = blob;
    ch.y = make_layout(L, A - 3u * L, T, nA);
    ch.use_alt = a.use_alt;
    const uint32_t rr = r0 - a.r_base;
    uint8_t* gdst = reinterpret_cast<uint8_t*>(a.xyz + 3u * a0);
    uint8_t* sdst = smem + so.o_out + ((uintptr_t)gdst & 15u);
    ch.out_xyz = reinterpret_cast<float*>(sdst);
    ch.out_type = nullptr; ch.out_bfac = nullptr; ch.out_meta = nullptr; ch.out_title = nullptr;
    ch.aoff = reinterpret_cast<uint32_t*>(smem + so.o_aoff);
    ch.segid = smem + so.o_segid;
    ch.tor = nullptr; ch.ang = nullptr;
    ch.order = reinterpret_cast<uint16_t*>(smem + so.o_order + 128u);
    ch.bins = reinterpret_cast<uint32_t*>(smem + so.o_order);
    ch.seg = reinterpret_cast<float*>(smem + so.o_seg);
    ch.rev = a.rev + 9u * (size_t)rr;
    ch.loc = a.loc + 9u * (size_t)rr;
    __builtin_assume(__isShared(ch.out_xyz));
    __builtin_assume(__isShared(ch.aoff));
    __builtin_assume(__isShared(ch.segid));
    __builtin_assume(__isShared(ch.seg));
    __builtin_assume(__isShared(ch.order));
    __builtin_assume(__isShared(ch.bins));
    // inputs of the first phases: asynchronous copies straight into shared memory, all in flight together;
    // the blend's inputs (written by the front kernel a whole batch ago, i.e. in HBM) are pulled towards L2 meanwhile
    {
        const char* pl = reinterpret_cast<const char*>(ch.loc);
        const char* pv = reinterpret_cast<const char*>(ch.rev);
        for (uint32_t off = cx.tid * 128u; off < 36u * L; off += cx.nthr * 128u) { prefetch_l2(pl + off); prefetch_l2(pv + off); }
    }
    {
        const uint32_t* g_aoff = a.aoff + rr + (c - a.c0);
        for (uint32_t r = cx.tid; r <= L; r += cx.nthr) cp_async4(ch.aoff + r, g_aoff + r);
        const float4* gseg = reinterpret_cast<const float4*>(a.seg + (size_t)(g0 - a.s_base) * FCZ_SEG_FLOATS);
        float4* sseg = reinterpret_cast<float4*>(ch.seg);
        for (uint32_t e = cx.tid; e < nA * (FCZ_SEG_FLOATS / 4u); e += cx.nthr) cp_async16(sseg + e, gseg + e);
        uint8_t* codes = smem + so.o_codes;
        const uint8_t* gtype = a.res_type + r0;  // written by the front kernel
        for (uint32_t r = cx.tid; r < L; r += cx.nthr) codes[r] = gtype[r];
        ch.codes = codes;
        // (staging the side-chain bytes here as well was measured: 2.5 KB more shared memory per block costs the fifth
        // block per SM, 0.260 -> 0.327 ms on the headline batch)
        ch.sc = nullptr;
        cp_async_wait_all();
    }
    __syncthreads();
    for (uint32_t s = cx.tid; s + 1u < nA; s += cx.nthr) {
        const float* sg = ch.seg + s * FCZ_SEG_FLOATS;
        for (uint32_t r = seg_a0(sg); r < seg_a1(sg); r++) ch.segid[r] = (uint8_t)s;
    }
    __syncthreads();
    cx.mark(13);
    dec_blend(cx, a.tables, ch);
    __syncthreads();
    cx.mark(11);
    dec_side(cx, a.tables, ch);
    cx.mark(12);
    copy_out(cx, gdst, sdst, 12u * A);  // (fences every thread's writes and synchronises the block first)
    cx.mark(14);
}

// Decode launch plan for device-planned batches.  Sub-batches of ~sub_res residues (equal chain counts) bound the
// workspace; inside a sub-batch chains are binned by length into FCZ_DEC_TIERS tiers, each launched on its own
// with shared memory sized by the tier's actual maxima, so a few long chains do not cost the short ones their
// occupancy.  Written to pinned host memory: [0] count, then per bound k (chain, residue, segment-slot prefix),
// then per (sub-batch, tier) five words: chains, max residues, max anchors, max blob bytes, max atoms.
// list[t * n + c0_k + i] = i-th chain of tier t in sub-batch k.  One thread per chain; counts and maxima through
// device atomics in dmax (all zero on entry and on exit), published by the last block to finish.
#define FCZ_DEC_TIERS 4
__host__ __device__ inline uint32_t dec_tier_of(uint32_t L) { return L <= 384u ? 0u : (L <= 768u ? 1u : (L <= 1536u ? 2u : 3u)); }

__global__ void __launch_bounds__(256) k_plan_chunks(uint32_t n, uint32_t sub_res, const uint32_t* res_off, const uint32_t* seg_off,
                                                     const uint64_t* atom_off, const uint64_t* blob_off, const int32_t* status,
                                                     uint32_t* list, uint32_t* dmax, uint32_t* out, const uint32_t* title_off, DecDesc* desc) {
    __shared__ uint32_t s_last;
    const uint64_t R = res_off[n];
    uint64_t per64 = n ? ((uint64_t)sub_res * n + (R ? R - 1 : 0)) / (R ? R : 1) : 1;
    if (per64 < 1) per64 = 1;
    uint64_t nch64 = n ? (n + per64 - 1) / per64 : 0;
    if (nch64 > FCZ_MAX_CHUNKS) { per64 = (n + FCZ_MAX_CHUNKS - 1) / FCZ_MAX_CHUNKS; nch64 = (n + per64 - 1) / per64; }
    const uint32_t per = (uint32_t)per64, nch = (uint32_t)nch64;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) out[0] = nch;
        for (uint32_t k = threadIdx.x; k <= nch; k += blockDim.x) {
            uint64_t c = (uint64_t)k * per;
            if (c > n) c = n;
            out[1 + 3 * k] = (uint32_t)c;
            out[2 + 3 * k] = res_off[c];
            out[3 + 3 * k] = seg_off[c];
        }
    }
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = c < n && status[c] == FCZ_OK;
    uint32_t k = 0, t = 0, m[4] = {0, 0, 0, 0};
    if (valid) {
        k = c / per;
        m[0] = res_off[c + 1] - res_off[c];
        m[1] = seg_off[c + 1] - seg_off[c];
        m[2] = (uint32_t)(blob_off[c + 1] - blob_off[c]);
        m[3] = (uint32_t)(atom_off[c + 1] - atom_off[c]);
        t = dec_tier_of(m[0]);
    }
    size_t li = 0;  // the chain's position in the tier lists
    // a block whose chains all sit in one sub-batch (the usual case) aggregates in shared memory first
    const uint32_t cf = blockIdx.x * blockDim.x, cl = (cf + blockDim.x < n ? cf + blockDim.x : n) - 1u;
    if (cf < n && cf / per == cl / per) {
        __shared__ uint32_t s_agg[FCZ_DEC_TIERS][6];  // count, four maxima, base
        if (threadIdx.x < FCZ_DEC_TIERS * 6) (&s_agg[0][0])[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t pos = 0;
        if (valid) {
            pos = atomicAdd(&s_agg[t][0], 1u);
            for (int j = 0; j < 4; j++) atomicMax(&s_agg[t][1 + j], m[j]);
        }
        __syncthreads();
        if (threadIdx.x < FCZ_DEC_TIERS && s_agg[threadIdx.x][0]) {
            uint32_t* d = dmax + 5u * (FCZ_DEC_TIERS * (cf / per) + threadIdx.x);
            s_agg[threadIdx.x][5] = atomicAdd(&d[0], s_agg[threadIdx.x][0]);
            for (int j = 0; j < 4; j++) atomicMax(&d[1 + j], s_agg[threadIdx.x][1 + j]);
        }
        __syncthreads();
        if (valid) li = (size_t)t * n + (size_t)k * per + s_agg[t][5] + pos;
    } else if (valid) {
        uint32_t* d = dmax + 5u * (FCZ_DEC_TIERS * k + t);
        const uint32_t pos = atomicAdd(&d[0], 1u);
        li = (size_t)t * n + (size_t)k * per + pos;
        for (int j = 0; j < 4; j++) atomicMax(&d[1 + j], m[j]);
    }
    if (valid) {
        list[li] = c;
        if (desc) {
            DecDesc q;
            q.c = c; q.r0 = res_off[c]; q.L = m[0]; q.A = m[3];
            q.t0 = title_off[c]; q.T = title_off[c + 1] - q.t0; q.g0 = seg_off[c]; q.nA = m[1];
            q.a0 = atom_off[c]; q.b0 = blob_off[c]; q.blob_len = m[2]; q.status = FCZ_OK;
            q.pad[0] = 0u; q.pad[1] = 0u;
            desc[li] = q;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&dmax[5 * FCZ_DEC_TIERS * FCZ_MAX_CHUNKS], 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last) {
        __threadfence();
        uint32_t* mx = out + 1 + 3 * (nch + 1);
        for (uint32_t i = threadIdx.x; i < 5u * FCZ_DEC_TIERS * nch; i += blockDim.x) mx[i] = atomicExch(&dmax[i], 0u);
        if (threadIdx.x == 0) dmax[5 * FCZ_DEC_TIERS * FCZ_MAX_CHUNKS] = 0u;
    }
}

// ============================================================================================= text
// PDB text of decoded chains (writeAtomCoordinatesToPDB, src/atom_coordinate.cpp:220-291) and the extract scans
// (Foldcomp::extract, src/foldcomp.cpp:1260-1336).  Formatting itself is fcz_text.h (shared with the CPU model).
//   k_pdb_plan  block per chain: residue -> first atom and residue -> first text byte (exact: over-long fields are
//               measured), the chain's text bytes and its number of emit units;
//   k_pdb_emit  block per unit of FCZ_PDB_UNIT_RES residues: thread per atom formats its 81-byte line into a
//               shared-memory image that has the 16-byte phase of its destination, then 128-bit stores.
// HBM-bound by the text written (81 B per atom against 12 B read).

struct PdbArgs {
    const uint32_t* res_off;
    const uint64_t* atom_off;
    const uint32_t* title_off;
    const uint8_t* res_type;
    const float* bfactor;
    const float* xyz;
    const char* titles;
    const fcz_chain_meta* meta;
    const TextTables* tt;
    int32_t use_alt;
    uint32_t n;
    uint32_t* aoff;          // [n_res + n] workspace: chain c at res_off[c] + c
    uint32_t* toff;          // same shape
    uint32_t* text_bytes;    // [n] plan output
    uint32_t* units;         // [n] plan output
    const uint64_t* text_off;  // [n+1] scans of the two
    const uint32_t* unit_off;
    uint32_t* unit_chain;    // [total units] chain of every emit unit (k_pdb_unit_map)
    char* text;
    uint32_t unit0;          // first unit of this launch (host-memory batches emit in slabs)
};

__device__ __forceinline__ PdbChain pdb_chain(const PdbArgs& a, uint32_t c) {
    PdbChain ch;
    const uint32_t r0 = a.res_off[c];
    const uint64_t a0 = a.atom_off[c];
    const uint32_t t0 = a.title_off[c];
    ch.L = a.res_off[c + 1] - r0;
    ch.A = (uint32_t)(a.atom_off[c + 1] - a0);
    ch.title_len = a.title_off[c + 1] - t0;
    ch.type = a.res_type + r0;
    ch.bfac = a.bfactor + r0;
    ch.X = a.xyz + 3u * a0;
    ch.title = a.titles + t0;
    ch.meta = a.meta + c;
    ch.use_alt = a.use_alt;
    ch.aoff = a.aoff + r0 + c;
    ch.toff = a.toff + r0 + c;
    return ch;
}

__global__ void __launch_bounds__(128) k_pdb_plan(PdbArgs a) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t scratch;
    const uint32_t c = blockIdx.x;
    DevCtx cx = block_ctx(wsum);
    const PdbChain ch = pdb_chain(a, c);
    const uint32_t total = pdb_plan_chain(cx, a.tt, ch, &scratch);
    if (threadIdx.x == 0) {
        a.text_bytes[c] = total;
        a.units[c] = (ch.L + FCZ_PDB_UNIT_RES - 1u) / FCZ_PDB_UNIT_RES;
    }
}

__global__ void __launch_bounds__(128) k_pdb_emit(PdbArgs a) {
    __shared__ __align__(16) char stage[FCZ_PDB_STAGE_BYTES];
    __shared__ uint32_t us[2 * FCZ_PDB_UNIT_RES + 2];
    const uint32_t u = a.unit0 + blockIdx.x;
    const uint32_t c = __ldg(a.unit_chain + u);
    DevCtx cx = block_ctx(nullptr);
    const PdbChain ch = pdb_chain(a, c);
    const uint32_t r_lo = (u - a.unit_off[c]) * FCZ_PDB_UNIT_RES;
    const uint32_t r_hi = r_lo + FCZ_PDB_UNIT_RES < ch.L ? r_lo + FCZ_PDB_UNIT_RES : ch.L;
    pdb_emit_unit(cx, a.tt, ch, r_lo, r_hi, a.text + a.text_off[c], stage, us);
}

// thread per chain: the chain index of each of its emit units (so that k_pdb_emit needs no search)
__global__ void __launch_bounds__(256) k_pdb_unit_map(PdbArgs a) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n) return;
    for (uint32_t u = a.unit_off[c]; u < a.unit_off[c + 1]; u++) a.unit_chain[u] = c;
}

struct ExtractArgs {
    const uint64_t* blob_off;
    const uint8_t* bytes;
    const TextTables* tt;
    int32_t type;
    uint32_t digits;
    uint32_t n;
    uint32_t* text_bytes;      // [n] plan output
    int32_t* status;           // [n]
    const uint64_t* text_off;  // [n+1]
    char* text;
};
// thread per blob: header check (Foldcomp::read, src/foldcomp.cpp:904-924) and output size
__global__ void __launch_bounds__(256) k_extract_plan(ExtractArgs a) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.n) return;
    const uint8_t* blob = a.bytes + a.blob_off[c];
    const uint64_t len = a.blob_off[c + 1] - a.blob_off[c];
    int st = FCZ_OK;
    uint32_t out = 0;
    if (len < HDR_BYTES || blob[0] != 'F' || blob[1] != 'C' || blob[2] != 'M' || blob[3] != 'P') st = FCZ_E_MAGIC;
    else {
        const Layout y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
        if ((uint64_t)get_u32(blob + OFF_LENTITLE) > len || (uint64_t)get_u32(blob + OFF_NSC) > len || (uint64_t)y.size > len) st = FCZ_E_TRUNCATED;
        else out = extract_len(y.L, a.type, a.digits);
    }
    a.text_bytes[c] = out;
    a.status[c] = st;
}
__global__ void __launch_bounds__(128) k_extract(ExtractArgs a) {
    const uint32_t c = blockIdx.x;
    if (a.status[c] != FCZ_OK) return;
    const uint8_t* blob = a.bytes + a.blob_off[c];
    const Layout y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
    DevCtx cx = block_ctx(nullptr);
    extract_chain(cx, a.tt, blob, y, a.type, a.digits, a.text + a.text_off[c]);
}

// Foldcomp::read + checkValidity (src/foldcomp.cpp:904-1036, 1492-1532) for every blob: one warp per blob scans the
// record, side-chain and B-factor sections for a non-zero entry (see fcz_check_batch in include/fcz_engine.h).
struct CheckArgs {
    const uint64_t* blob_off;
    const uint8_t* bytes;
    uint32_t n;
    int32_t* read_status;  // may be null
    int32_t* validity;     // may be null
};
__global__ void __launch_bounds__(256) k_check(CheckArgs a) {
    const uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31u;
    if (c >= a.n) return;
    const uint8_t* blob = a.bytes + a.blob_off[c];
    const uint64_t len = a.blob_off[c + 1] - a.blob_off[c];
    int st = FCZ_OK, v = FCZ_V_SUCCESS;
    if (len < 4u || blob[0] != 'F' || blob[1] != 'C' || blob[2] != 'M' || blob[3] != 'P') st = FCZ_E_MAGIC;
    else if (len < HDR_BYTES) { st = FCZ_E_TRUNCATED; v = FCZ_V_BACKBONE_COUNT_MISMATCH; }
    else {
        const Layout y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
        // sections in file order; 64-bit sums: lenTitle and nSideChainTorsion are 32-bit fields of an untrusted header
        const uint64_t e_rec = (uint64_t)HDR_BYTES + 40ull * y.n_anchor + y.title_len + 13ull + 8ull * y.L;
        const uint64_t e_sc = e_rec + y.n_sc, e_tmp = e_sc + 8ull + y.L;
        if (e_rec > len) { st = FCZ_E_TRUNCATED; v = FCZ_V_BACKBONE_COUNT_MISMATCH; }
        else if (e_sc > len) { st = FCZ_E_TRUNCATED; v = FCZ_V_SIDECHAIN_COUNT_MISMATCH; }
        else if (e_tmp > len) { st = FCZ_E_TRUNCATED; v = FCZ_V_TEMP_FACTOR_COUNT_MISMATCH; }
        else {
            uint32_t any_bb = 0, any_sc = 0, any_t = 0;
            const uint8_t* rec = blob + y.o_rec;
            // phi | psi | omega non-zero <=> (byte0 & 7) | byte1 | byte2 | byte3 | byte4 (convertBytesToBackboneChain, 60-77)
            for (uint32_t r = lane; r < y.L; r += 32u) {
                const uint8_t* p = rec + 8u * r;
                any_bb |= (uint32_t)(p[0] & 7u) | p[1] | p[2] | p[3] | p[4];
            }
            for (uint32_t i = lane; i < y.n_sc; i += 32u) any_sc |= blob[y.o_sc + i];
            for (uint32_t i = lane; i < y.L; i += 32u) any_t |= blob[y.o_temp + 8u + i];
            any_bb = __any_sync(0xffffffffu, any_bb != 0u);
            any_sc = __any_sync(0xffffffffu, any_sc != 0u);
            any_t = __any_sync(0xffffffffu, any_t != 0u);
            v = !any_bb ? FCZ_V_EMPTY_BACKBONE_ANGLE : (!any_sc ? FCZ_V_EMPTY_SIDECHAIN_ANGLE : (!any_t ? FCZ_V_EMPTY_TEMP_FACTOR : FCZ_V_SUCCESS));
        }
    }
    if (lane == 0) {
        if (a.read_status) a.read_status[c] = st;
        if (a.validity) a.validity[c] = v;
    }
}

// Continuised backbone angles of every residue record (decompressBackboneChain, src/foldcomp.cpp:122-153 with
// _continuize 155-158: q * cont_f + min, two float roundings): six floats per residue in header order
// phi, psi, omega, N-CA-C, CA-C-N, C-N-CA.  What Foldcomp::decompress leaves in its phi / psi / omega / *_angle members
// (src/foldcomp.cpp:783-804) and the CPython get_data() returns (foldcomp/foldcomp.cxx:497-560).  Thread per residue.
struct AnglesArgs {
    const uint64_t* blob_off;
    const uint8_t* bytes;
    const int32_t* status;
    const uint64_t* res_off;  // [n+1] scan of the residue counts
    float* angles;            // [6 * residues]
    uint32_t n;
};
__global__ void __launch_bounds__(128) k_unpack_angles(AnglesArgs a) {
    const uint32_t c = blockIdx.x;
    if (a.status[c] != FCZ_OK) return;
    const uint8_t* blob = a.bytes + a.blob_off[c];
    const Layout y = make_layout(get_u16(blob + OFF_NRES), get_u32(blob + OFF_NSC), get_u32(blob + OFF_LENTITLE), blob[OFF_NANCHOR]);
    float mins[6], cfs[6];
    for (int k = 0; k < 6; k++) { mins[k] = get_f32(blob + OFF_MINS + 4 * k); cfs[k] = get_f32(blob + OFF_CONTFS + 4 * k); }
    float* out = a.angles + 6u * a.res_off[c];
    for (uint32_t r = threadIdx.x; r < y.L; r += blockDim.x) {
        const Record q = unpack_record(blob + y.o_rec + 8u * r);
        out[6u * r + 0u] = continuize(q.phi, mins[A_PHI], cfs[A_PHI]);
        out[6u * r + 1u] = continuize(q.psi, mins[A_PSI], cfs[A_PSI]);
        out[6u * r + 2u] = continuize(q.omg, mins[A_OMEGA], cfs[A_OMEGA]);
        out[6u * r + 3u] = continuize(q.nca, mins[A_NCAC], cfs[A_NCAC]);
        out[6u * r + 4u] = continuize(q.cac, mins[A_CACN], cfs[A_CACN]);
        out[6u * r + 5u] = continuize(q.cnc, mins[A_CNCA], cfs[A_CNCA]);
    }
}

// Backbone angles before quantisation (enc_raw_angles): block per chain; the residue -> atom offsets are scanned into a
// global workspace first (any chain length), then one thread per value.
struct RawAnglesArgs {
    const uint32_t* res_off;
    const uint64_t* atom_off;
    const uint8_t* res_type;
    const float* xyz;
    const Tables* tables;
    uint32_t* aoff;   // [n_res + n] workspace
    float* angles;    // [6 * n_res]
    uint32_t n;
};
__global__ void __launch_bounds__(128) k_raw_angles(RawAnglesArgs a) {
    __shared__ uint32_t wsum[32];
    const uint32_t c = blockIdx.x;
    DevCtx cx = block_ctx(wsum);
    const uint32_t r0 = a.res_off[c], L = a.res_off[c + 1] - r0;
    if (L == 0u) return;
    uint32_t* aoff = a.aoff + r0 + c;
    const uint8_t* type = a.res_type + r0;
    const uint32_t chunk = (L + blockDim.x - 1u) / blockDim.x;
    uint32_t b0 = threadIdx.x * chunk; if (b0 > L) b0 = L;
    uint32_t b1 = b0 + chunk; if (b1 > L) b1 = L;
    uint32_t sum = 0;
    for (uint32_t r = b0; r < b1; r++) sum += natoms_packed(type[r] & 31u);
    uint32_t base = cx.excl_scan(sum);
    for (uint32_t r = b0; r < b1; r++) { aoff[r] = base; base += natoms_packed(type[r] & 31u); }
    if (b1 == L && b0 < L) aoff[L] = base;
    __syncthreads();
    EncChain ch;
    memset(&ch, 0, sizeof ch);
    ch.L = L; ch.X = a.xyz + 3ull * a.atom_off[c]; ch.aoff = aoff; ch.type = type;
    enc_raw_angles(cx, ch, a.angles + 6ull * r0);
}

// ============================================================================================ PDB text in
// fcz_parse.h on the GPU: k_parse_lines counts the lines of every entry (sizes the workspace), k_parse_plan (block per
// entry) finds and parses the ATOM records, drops alternative positions and splits residues, k_parse_emit (block per
// entry) fills the canonical layout.
struct ParseArgs {
    const uint64_t* text_off;
    const char* text;
    uint32_t n;
    const ParseTables* pt;
    uint32_t* n_lines;          // [n] lines per entry (k_parse_lines)
    const uint64_t* line_off;   // [n+1] their scan: workspace slots
    uint32_t* lines;            // [total_lines + n]
    uint32_t* rstart;           // [total_lines + n]
    RawAtom* raw;               // [2 * total_lines]
    uint32_t* scratch;          // [8 * n]
    uint32_t* v_res;            // [n] residues per entry (plan output)
    uint32_t* v_atoms;          // [n] table slots per entry
    int32_t* status;            // [n]
    const uint32_t* res_off;    // [n+1]
    const uint64_t* atom_off;   // [n+1]
    uint8_t* res_type;
    float* bfactor;
    float* xyz;
    fcz_chain_meta* meta;
};
__global__ void __launch_bounds__(256) k_parse_lines(ParseArgs a) {
    __shared__ uint32_t s_cnt;
    const uint32_t c = blockIdx.x;
    const char* t = a.text + a.text_off[c];
    const uint64_t len = a.text_off[c + 1] - a.text_off[c];
    if (threadIdx.x == 0) s_cnt = 0u;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) cnt += t[i] == '\n';
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) a.n_lines[c] = s_cnt + ((len && t[len - 1] != '\n') ? 1u : 0u);
}
__device__ __forceinline__ ParseEntry parse_entry_of(const ParseArgs& a, uint32_t c) {
    ParseEntry e;
    e.text = a.text + a.text_off[c];
    e.len = (uint32_t)(a.text_off[c + 1] - a.text_off[c]);
    const uint64_t l0 = a.line_off[c];
    e.max_lines = (uint32_t)(a.line_off[c + 1] - l0);
    e.lines = a.lines + l0 + c;
    e.rstart = a.rstart + l0 + c;
    e.raw = a.raw + 2ull * l0;
    e.scratch = a.scratch + 8ull * c;
    return e;
}
__global__ void __launch_bounds__(256) k_parse_plan(ParseArgs a) {
    __shared__ uint32_t wsum[32];
    const uint32_t c = blockIdx.x;
    DevCtx cx = block_ctx(wsum);
    const ParseEntry e = parse_entry_of(a, c);
    if (e.max_lines == 0u || (a.text_off[c + 1] - a.text_off[c]) >= 0xFFFFFFFFull) {  // empty text (or one beyond 4 GB)
        if (threadIdx.x == 0) { a.v_res[c] = 0u; a.v_atoms[c] = 0u; a.status[c] = FCZ_E_PARSE_NOATOM; e.scratch[PS_FLAG] = 1u; }
        return;
    }
    parse_entry_plan(cx, a.pt, e);
    if (threadIdx.x == 0) {
        const uint32_t f = e.scratch[PS_FLAG];
        a.status[c] = f == 0u ? FCZ_OK : (f == 1u ? FCZ_E_PARSE_NOATOM : (f == 2u ? FCZ_E_PARSE_CHAINS : (f == 3u ? FCZ_E_PARSE_RECORD : (f == 4u ? FCZ_E_PARSE_NUMBER : FCZ_E_PARSE_GAPS))));
        a.v_res[c] = f ? 0u : e.scratch[PS_NRES];
        a.v_atoms[c] = f ? 0u : e.scratch[PS_NSLOT];
    }
}
__global__ void __launch_bounds__(256) k_parse_emit(ParseArgs a) {
    __shared__ uint32_t wsum[32];
    const uint32_t c = blockIdx.x;
    if (a.status[c] != FCZ_OK) return;
    DevCtx cx = block_ctx(wsum);
    const ParseEntry e = parse_entry_of(a, c);
    parse_entry_emit(cx, a.pt, e, a.res_type + a.res_off[c], a.bfactor + a.res_off[c], a.xyz + 3ull * a.atom_off[c], a.meta + c);
}

// ============================================================================================ scans
// Exclusive scans of up to three per-chain u32 arrays into offsets (u32/u64), tile = 2048 chains.

#define SCAN_TILE 2048
#define SCAN_THREADS 256
#define SCAN_ITEMS (SCAN_TILE / SCAN_THREADS)

struct ScanArgs {
    uint32_t n;
    int narr;
    const uint32_t* in[4];
    void* out[4];       // [n+1]
    int out64[4];       // 1: uint64 output, 0: uint32
    uint64_t* partial;  // [ntiles*4]
    uint64_t* totals;   // [4]
};

__device__ __forceinline__ uint64_t block_sum64(uint64_t v, uint64_t* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    uint64_t s = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += sh[w];
    __syncthreads();
    return s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partials(ScanArgs a) {
    __shared__ uint64_t sh[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    for (int j = 0; j < a.narr; j++) {
        uint64_t s = 0;
        for (int i = 0; i < SCAN_ITEMS; i++)
            if (base + i < a.n) s += a.in[j][base + i];
        s = block_sum64(s, sh);
        if (threadIdx.x == 0) a.partial[blockIdx.x * 4 + j] = s;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(ScanArgs a) {
    __shared__ uint64_t sh[32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = 0; j < a.narr; j++) {
        uint64_t pre = 0;
        for (uint32_t t = threadIdx.x; t < blockIdx.x; t += blockDim.x) pre += a.partial[t * 4 + j];
        pre = block_sum64(pre, sh);
        uint32_t v[SCAN_ITEMS];
        uint64_t s = 0;
        for (int i = 0; i < SCAN_ITEMS; i++) {
            v[i] = (base + i < a.n) ? a.in[j][base + i] : 0u;
            s += v[i];
        }
        uint64_t x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t yv = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += yv;
        }
        if (lane == 31) sh[warp] = x;
        __syncthreads();
        uint64_t wb = 0;
        for (int w = 0; w < warp; w++) wb += sh[w];
        uint64_t tile_total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) tile_total += sh[w];
        __syncthreads();
        uint64_t off = pre + wb + x - s;
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (base + i < a.n) {
                if (a.out64[j]) ((uint64_t*)a.out[j])[base + i] = off;
                else ((uint32_t*)a.out[j])[base + i] = (uint32_t)off;
            }
            off += v[i];
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
            uint64_t tot = pre + tile_total;
            if (a.out64[j]) ((uint64_t*)a.out[j])[a.n] = tot;
            else ((uint32_t*)a.out[j])[a.n] = (uint32_t)tot;
            a.totals[j] = tot;
        }
    }
}

// ============================================================================================ engine

#define FCZ_HOST_CHUNKS 32u  // chunks of a host-memory batch
// Transfer shape of a host-memory batch, chosen from tools/e2e_sweep.py on a B200 (profiles/r01_v12_e2e_sweep.txt): two
// engines (an encode next to a decode) reach 346 M residues/s with ONE coordinate copy per call of <= 96 MB, 305 M with
// 12 MB chunks queued at once and 275-299 M with 12 MB chunks and a bounded queue.  Both knobs stay as environment
// overrides (FCZ_CHUNK_MB, FCZ_H2D_QUEUE, FCZ_D2H_QUEUE); a depth of 0 queues every chunk at once.
#define FCZ_CHUNK_BYTES (96ull << 20)
#define FCZ_H2D_DEPTH 0u     // input chunks queued ahead on the copy engine (encode_host)
#define FCZ_D2H_DEPTH 0u     // coordinate chunks queued ahead on the D2H copy engine (decode_host)
#define FCZ_D2H_GROUP 4u     // encode: chunks whose blobs come back in one copy
#define FCZ_DEC_H2D_PIECES 2u  // copies the blobs of a host-memory decode go up in (see decode_host)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct fcz_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the pipelined host-memory path
    cudaStream_t s_tier[8] = {};  // length tiers (encode: FCZ_NTIER, decode: FCZ_DEC_TIERS) run side by side
    cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
    std::vector<cudaEvent_t> ev_pool;              // events of that path, reused across calls
    fcz_opts opts;
    int num_sms = 148;
    TierCfg enc_tier[FCZ_NTIER];
    int enc_occ[FCZ_NTIER];
    uint32_t enc_prefetch_next = 0;  // FCZ_ENC_PREFETCH_NEXT=1: pull the next chain towards L2 a trip ahead (measured 4 % SLOWER on the headline batch)
    uint32_t enc_pad = 0;            // FCZ_ENC_PAD_KB: extra dynamic shared memory per block, i.e. fewer blocks per SM (A/B)
    Tables* d_tables = nullptr;
    TextTables* d_text_tables = nullptr;
    // text emitter (fcz_pdb_text_plan -> fcz_pdb_text_batch)
    DevBuf ws_aoff, ws_toff, d_unit_off, d_unit_chain, d_text_off, d_text;
    struct PdbPlan { uint32_t n = 0; uint64_t total_bytes = 0; uint32_t total_units = 0; bool valid = false; } pdb;
    std::vector<uint32_t> h_unit_off;
    // PDB text parser (fcz_parse_pdb_plan -> fcz_parse_pdb_batch): tables, per-line workspace
    ParseTables* d_parse_tables = nullptr;
    DevBuf p_line_off, p_lines, p_rstart, p_raw, p_scratch;
    struct ParsePlan { uint32_t n = 0; uint64_t n_res = 0, n_atoms = 0; int32_t* status = nullptr; bool valid = false; } parse;
    std::vector<int32_t> h_parse_status;
    // plan scratch
    DevBuf v0, v1, v2, v3, status, tier_list, partial;
    uint32_t* d_counters = nullptr;  // [FCZ_NTIER] counts, [FCZ_NTIER] tickets
    uint64_t* d_totals = nullptr;    // [3]
    uint32_t* h_counters = nullptr;  // pinned mirror
    uint64_t* h_totals = nullptr;
    uint64_t chunk_bytes = FCZ_CHUNK_BYTES;  // coordinates per chunk of a host-memory batch (FCZ_CHUNK_MB)
    uint32_t h2d_depth = FCZ_H2D_DEPTH, d2h_depth = FCZ_D2H_DEPTH;  // FCZ_H2D_QUEUE / FCZ_D2H_QUEUE, 0 = unlimited
    bool dec_desc_valid = false;  // the last decode plan filled d_dec_desc (device-planned batches)
    uint32_t dec_sub_res = FCZ_SUB_RESIDUES;  // residues per decode sub-batch (FCZ_DEC_SUB_RESIDUES overrides)
    uint32_t* h_bounds = nullptr;    // pinned: [0] nchunks, then chain / residue / segment-slot bounds of the decode sub-batches
    // staging for host-memory batches
    DevBuf d_res_off, d_atom_off, d_title_off, d_res_type, d_bfactor, d_xyz, d_titles, d_meta, d_blob_off, d_bytes, d_status;
    DevBuf d_list, d_tickets, enc_gws, d_stage;
    void* h_stage = nullptr;  // pinned staging for the per-chain arrays of a host-memory decode (one H2D copy)
    size_t h_stage_cap = 0;
    struct { uint64_t *blob_off, *atom_off; uint32_t *res_off, *title_off, *seg_off, *list; int32_t* status; } dh = {};
    DevBuf d_seg_off, sc_aoff, sc_segid, sc_tor, sc_ang, sc_rev, sc_seg, sc_loc, d_submax, d_dec_list, d_dec_desc, d_enc_desc;  // decoder: segment offsets + hand-over workspace between the phase kernels
    // plan made on the host by fcz_decode_plan(host) for the following fcz_decode_batch(host)
    struct Launch { uint32_t chunk, tier, first, count; };
    struct HostPlan {
        uint32_t n = 0;
        std::vector<uint32_t> chunk_c0;   // [nchunks+1] chain ranges
        std::vector<uint32_t> list;       // chains grouped by (chunk, tier)
        std::vector<Launch> launches;
        std::vector<int32_t> status;
        std::vector<uint32_t> seg_off;    // decode: [n+1] prefix of anchors (segment-scratch slots)
    } hplan;
    uint64_t launches = 0;
    // optional per-kernel event timing
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;            // recorded since the last get_profile
    std::vector<cudaEvent_t> free_events;
    char err[512];
};

static cudaEvent_t take_event(fcz_engine* e) {
    cudaEvent_t ev = nullptr;
    if (!e->free_events.empty()) { ev = e->free_events.back(); e->free_events.pop_back(); }
    else cudaEventCreate(&ev);
    return ev;
}
struct ProfSpan {  // RAII: events around one kernel launch (or one fork-to-join span) when profiling is on
    fcz_engine* e; cudaEvent_t a = nullptr, b = nullptr; int kind; cudaStream_t st;
    ProfSpan(fcz_engine* e_, int kind_, cudaStream_t st_ = nullptr) : e(e_), kind(kind_), st(st_ ? st_ : e_->stream) {
        if (e->profiling) { a = take_event(e); b = take_event(e); cudaEventRecord(a, st); }
    }
    ~ProfSpan() {
        if (a) { cudaEventRecord(b, st); e->spans.push_back({a, b, kind}); }
    }
};

static int fail(fcz_engine* e, int code, const char* fmt, ...) {
    if (e) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(e->err, sizeof e->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t _e = (call);                                                                          \
        if (_e != cudaSuccess) return fail(e, FCZ_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

static int ensure(fcz_engine* e, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return FCZ_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    size_t cap = bytes + bytes / 8 + 256;
    CK(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return FCZ_OK;
}

extern "C" {

fcz_engine* fcz_engine_create(int device, const fcz_opts* opts) {
    fcz_engine* e = new (std::nothrow) fcz_engine();
    if (!e) return nullptr;
    e->err[0] = 0;
    e->device = device;
    e->opts.anchor_threshold = 25;
    e->opts.use_alt_atom_order = 0;
    e->opts.stream = nullptr;
    e->opts.terminate_blobs = 0;
    if (opts) e->opts = *opts;
    if (e->opts.anchor_threshold < 1) e->opts.anchor_threshold = 25;
    if (cudaSetDevice(device) != cudaSuccess) { delete e; return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; return nullptr; }
    e->num_sms = prop.multiProcessorCount;
    if (e->opts.stream) {
        e->stream = (cudaStream_t)e->opts.stream;
    } else {
        if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; return nullptr; }
        e->own_stream = true;
    }
    make_tiers(e->enc_tier);
    // experiment knobs (A/B runs, tools/kernel_ab.py): FCZ_ENC_XGLOBAL=1 reads coordinates from global memory instead of staging
    // them (smaller footprint, more CTAs per SM); FCZ_ENC_THREADS fixes the block size of the staged tiers
    {
        // Coordinates are read from global memory by default: without the 12 bytes per atom of staging a 352-residue
        // block needs 26.5 KB of shared memory instead of 71 KB, eight blocks share an SM instead of three, and the
        // measured kernel time is lower; FCZ_ENC_XGLOBAL=0 restores the bulk-copy staging.
        const char* v = getenv("FCZ_ENC_XGLOBAL");
        if (!v || atoi(v)) for (int i = 0; i < FCZ_NTIER; i++) { e->enc_tier[i].stage_x = 0; e->enc_tier[i].smem = enc_smem(e->enc_tier[i]).total; }
        if (const char* q = getenv("FCZ_ENC_PREFETCH_NEXT")) e->enc_prefetch_next = atoi(q) ? 1u : 0u;
        if (const char* q = getenv("FCZ_ENC_PAD_KB")) {
            e->enc_pad = (uint32_t)atoi(q) * 1024u;
            for (int i = 0; i < FCZ_NTIER - 2; i++) if (e->enc_tier[i].smem + e->enc_pad <= 227u * 1024u) e->enc_tier[i].smem += e->enc_pad;
        }
    }
    bool ok = true;
    ok &= cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking) == cudaSuccess;
    ok &= cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 8; i++) {
        ok &= cudaStreamCreateWithFlags(&e->s_tier[i], cudaStreamNonBlocking) == cudaSuccess;
        ok &= cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    }
    ok &= cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < FCZ_NTIER; i++) {
        if (e->enc_tier[i].smem > 227u * 1024u) {
            fprintf(stderr, "fcz_engine_create: tier %d needs %u bytes of shared memory (> 227 KB)\n", i, e->enc_tier[i].smem);
            ok = false;
        }
    }
    for (int i = 0; i < FCZ_NTIER && ok; i++) {
        ok &= cudaFuncSetAttribute(k_encode, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_encode_long, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_dec_front, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_dec_back, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        ok &= cudaFuncSetAttribute(k_dec_stitch_t, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        if (const char* v = getenv("FCZ_DEC_SUB_RESIDUES")) { long q = atol(v); if (q > 0) e->dec_sub_res = (uint32_t)q; }
        if (const char* v = getenv("FCZ_CHUNK_MB")) { long q = atol(v); if (q > 0) e->chunk_bytes = (uint64_t)q << 20; }
        if (const char* v = getenv("FCZ_H2D_QUEUE")) e->h2d_depth = (uint32_t)atol(v);
        if (const char* v = getenv("FCZ_D2H_QUEUE")) e->d2h_depth = (uint32_t)atol(v);
        // 64 registers per thread = 1024 threads per SM, shared between the CTAs the shared-memory footprint lets in
        int occ = 0;
        ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode, 128, e->enc_tier[i].smem) == cudaSuccess;
        if (occ < 1) occ = 1;
        // block size: as many threads as the register file leaves each of the `occ` blocks (whole warps)
        uint32_t lim = (FCZ_ENC_MAXREG <= 64 ? 1024u : 65536u / FCZ_ENC_MAXREG) / (uint32_t)occ;
        uint32_t thr = lim & ~31u;
        if (thr < 96u) thr = 96u;
        if (const char* v = getenv("FCZ_ENC_THREADS")) { const long q = atol(v); if (q >= 32 && e->enc_tier[i].staged) thr = (uint32_t)q; }
        if (thr > 768u) thr = 768u;
        e->enc_tier[i].threads = thr < 64u ? 64u : thr;
        ok &= cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_encode, (int)e->enc_tier[i].threads, e->enc_tier[i].smem) == cudaSuccess;
        e->enc_occ[i] = occ > 0 ? occ : 1;
        if (e->enc_tier[i].gws) { e->enc_tier[i].threads = 768u; e->enc_occ[i] = 1; }  // one wide block per SM walks a long chain (24 warps: FCZ_RED_FLOATS(24))
    }
    Tables h;
    build_tables(&h);
    TextTables ht;
    build_text_tables(&ht);
    ok &= cudaMalloc(&e->d_text_tables, sizeof(TextTables)) == cudaSuccess;
    if (ok) ok &= cudaMemcpy(e->d_text_tables, &ht, sizeof(TextTables), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMalloc(&e->d_tables, sizeof(Tables)) == cudaSuccess;
    ok &= cudaMalloc(&e->d_counters, sizeof(uint32_t) * (2 * FCZ_NTIER + 2)) == cudaSuccess;
    ok &= cudaMalloc(&e->d_totals, sizeof(uint64_t) * 4) == cudaSuccess;
    ok &= cudaMallocHost(&e->h_counters, sizeof(uint32_t) * (2 * FCZ_NTIER + 2)) == cudaSuccess;
    ok &= cudaMallocHost(&e->h_totals, sizeof(uint64_t) * 4) == cudaSuccess;
    ok &= cudaMallocHost(&e->h_bounds, sizeof(uint32_t) * ((3 + 5 * FCZ_DEC_TIERS) * (FCZ_MAX_CHUNKS + 1) + 1)) == cudaSuccess;
    if (ok) ok &= cudaMemcpy(e->d_tables, &h, sizeof(Tables), cudaMemcpyHostToDevice) == cudaSuccess;
    ok &= cudaMalloc(&e->d_submax.p, sizeof(uint32_t) * (5 * FCZ_DEC_TIERS * FCZ_MAX_CHUNKS + 1)) == cudaSuccess;  // k_plan_chunks keeps it zero
    if (ok) { e->d_submax.cap = sizeof(uint32_t) * (5 * FCZ_DEC_TIERS * FCZ_MAX_CHUNKS + 1); ok &= cudaMemset(e->d_submax.p, 0, e->d_submax.cap) == cudaSuccess; }
    if (!ok) {
        fprintf(stderr, "fcz_engine_create: %s\n", cudaGetErrorString(cudaGetLastError()));
        delete e;
        return nullptr;
    }
    return e;
}

void fcz_engine_destroy(fcz_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    DevBuf* bufs[] = {&e->v0, &e->v1, &e->v2, &e->v3, &e->status, &e->tier_list, &e->partial, &e->d_res_off, &e->d_atom_off,
                      &e->d_title_off, &e->d_res_type, &e->d_bfactor, &e->d_xyz, &e->d_titles, &e->d_meta,
                      &e->d_blob_off, &e->d_bytes, &e->d_status, &e->d_list, &e->d_tickets, &e->enc_gws, &e->d_stage, &e->ws_aoff, &e->ws_toff, &e->d_unit_off, &e->d_unit_chain, &e->d_text_off, &e->d_text,
                      &e->d_seg_off, &e->sc_aoff, &e->sc_segid, &e->sc_tor, &e->sc_ang, &e->sc_rev, &e->sc_seg, &e->sc_loc, &e->d_submax, &e->d_dec_list, &e->d_dec_desc, &e->d_enc_desc,
                      &e->p_line_off, &e->p_lines, &e->p_rstart, &e->p_raw, &e->p_scratch};
    for (DevBuf* b : bufs)
        if (b->p) cudaFree(b->p);
    if (e->d_tables) cudaFree(e->d_tables);
    if (e->d_text_tables) cudaFree(e->d_text_tables);
    if (e->d_parse_tables) cudaFree(e->d_parse_tables);
    if (e->d_counters) cudaFree(e->d_counters);
    if (e->d_totals) cudaFree(e->d_totals);
    if (e->h_counters) cudaFreeHost(e->h_counters);
    if (e->h_totals) cudaFreeHost(e->h_totals);
    if (e->h_bounds) cudaFreeHost(e->h_bounds);
    if (e->h_stage) cudaFreeHost(e->h_stage);
    for (auto& s : e->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto ev : e->free_events) cudaEventDestroy(ev);
    for (auto ev : e->ev_pool) cudaEventDestroy(ev);
    for (int i = 0; i < 8; i++) {
        if (e->s_tier[i]) cudaStreamDestroy(e->s_tier[i]);
        if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
    }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

int fcz_engine_set_opts(fcz_engine* e, const fcz_opts* opts) {
    if (!e || !opts) return FCZ_E_ARG;
    if (opts->anchor_threshold < 1) return fail(e, FCZ_E_ARG, "anchor_threshold must be >= 1");
    e->opts.anchor_threshold = opts->anchor_threshold;
    e->opts.use_alt_atom_order = opts->use_alt_atom_order;
    e->opts.terminate_blobs = opts->terminate_blobs;
    if (opts->stream != e->opts.stream) {
        if (e->own_stream) {
            cudaStreamSynchronize(e->stream);
            cudaStreamDestroy(e->stream);
            e->own_stream = false;
        }
        if (opts->stream) {
            e->stream = (cudaStream_t)opts->stream;
        } else {
            if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(e, FCZ_E_CUDA, "stream");
            e->own_stream = true;
        }
        e->opts.stream = opts->stream;
    }
    return FCZ_OK;
}

uint64_t fcz_encode_bound(uint64_t n_chains, uint64_t n_res, uint64_t n_atoms, uint64_t n_title_bytes, int32_t b) {
    if (b < 1) b = 1;
    return 98ull * n_chains + 40ull * (n_res / (uint64_t)b + 2ull * n_chains) + n_title_bytes + 6ull * n_res + n_atoms;  // 97 + a terminator
}

void* fcz_host_alloc(size_t bytes) {
    void* p = nullptr;
    return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void fcz_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int fcz_engine_sync(fcz_engine* e) {
    if (!e) return FCZ_E_ARG;
    CK(cudaStreamSynchronize(e->stream));
    return FCZ_OK;
}

uint64_t fcz_engine_launch_count(const fcz_engine* e) { return e ? e->launches : 0; }

#ifdef FCZ_PHASE_TIMING
// debug-only export: out[0..15] cycles, out[16..31] counts; resets the counters
int fcz_debug_phase_cycles(fcz_engine* e, unsigned long long* out) {
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof(unsigned long long) * 16));
    CK(cudaMemcpyFromSymbol(out + 16, g_phase_count, sizeof(unsigned long long) * 16));
    unsigned long long z[16] = {0};
    CK(cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z));
    CK(cudaMemcpyToSymbol(g_phase_count, z, sizeof z));
    return FCZ_OK;
}
#endif

int fcz_engine_set_profiling(fcz_engine* e, int enabled) {
    if (!e) return FCZ_E_ARG;
    e->profiling = enabled != 0;
    return FCZ_OK;
}

int fcz_engine_get_profile(fcz_engine* e, fcz_profile* out) {
    if (!e || !out) return FCZ_E_ARG;
    CK(cudaStreamSynchronize(e->stream));
    memset(out, 0, sizeof *out);
    for (auto& s : e->spans) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, s.a, s.b));
        if (s.kind == FCZ_PROF_ENCODE) { out->encode_kernel_ms += ms; out->encode_launches++; }
        else if (s.kind == FCZ_PROF_DECODE) { out->decode_kernel_ms += ms; out->decode_launches++; }
        if (s.kind >= 0 && s.kind < FCZ_PROF_KINDS) { out->kernel_ms[s.kind] += ms; out->kernel_launches[s.kind]++; }
        e->free_events.push_back(s.a);
        e->free_events.push_back(s.b);
    }
    e->spans.clear();
    return FCZ_OK;
}

const char* fcz_strerror(int code) {
    switch (code) {
        case FCZ_OK: return "ok";
        case FCZ_E_MAGIC: return "not an FCZ blob (bad magic)";
        case FCZ_E_TRUNCATED: return "blob truncated or inconsistent with its header";
        case FCZ_E_RESIDUE: return "residue code outside the 20 amino acids + UNK";
        case FCZ_E_LIMIT: return "chain exceeds a format or engine limit";
        case FCZ_E_CAPACITY: return "output buffer too small";
        case FCZ_E_CUDA: return "CUDA error";
        case FCZ_E_ARG: return "bad argument";
        case FCZ_E_PARSE_NOATOM: return "No ATOM lines found";
        case FCZ_E_PARSE_CHAINS: return "Multiple chains found";
        case FCZ_E_PARSE_RECORD: return "Malformed ATOM record";
        case FCZ_E_PARSE_NUMBER: return "numeric field outside the fixed-point grammar";
        case FCZ_E_PARSE_GAPS: return "discontinuous residue numbering (the chain splits into fragments)";
        default: return "unknown error";
    }
}
const char* fcz_last_error(const fcz_engine* e) { return e ? e->err : "no engine"; }

int fcz_type_natoms(int c) { return (c >= 0 && c < FCZ_NUM_CODES) ? FCZ_NATOMS[c] : 0; }
const char* fcz_type_name3(int c) { return (c >= 0 && c < FCZ_NUM_CODES) ? FCZ_NAME3[c] : ""; }
const char* fcz_type_atom_name(int c, int k) {
    return (c >= 0 && c < FCZ_NUM_CODES && k >= 0 && k < FCZ_MAX_ATOMS) ? FCZ_ATOM_NAME[c][k] : "";
}
int fcz_type_alt_slot(int c, int k) { return (c >= 0 && c < FCZ_NUM_CODES && k >= 0 && k < FCZ_MAX_ATOMS) ? FCZ_ALT[c][k] : 0; }
int fcz_type_pred(int c, int k, int w) {
    if (c < 0 || c >= FCZ_NUM_CODES || k < 0 || k >= FCZ_MAX_ATOMS || w < 0 || w > 2) return 0;
    return (FCZ_PRED[c][k] >> (4 * w)) & 15;
}
float fcz_type_bond_length(int c, int k) { return (c >= 0 && c < FCZ_NUM_CODES && k >= 0 && k < FCZ_MAX_ATOMS) ? FCZ_BLEN[c][k] : 0.f; }
float fcz_type_bond_angle(int c, int k) { return (c >= 0 && c < FCZ_NUM_CODES && k >= 0 && k < FCZ_MAX_ATOMS) ? FCZ_BANG[c][k] : 0.f; }

}  // extern "C"

// ------------------------------------------------------------------------------------ plan helpers

static int plan_buffers(fcz_engine* e, uint32_t n) {
    int rc;
    if ((rc = ensure(e, e->v0, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->v1, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->v2, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->v3, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->status, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->tier_list, 4ull * n * FCZ_NTIER + 4))) return rc;
    uint32_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if ((rc = ensure(e, e->partial, 32ull * (ntiles + 1)))) return rc;
    CK(cudaMemsetAsync(e->d_counters, 0, sizeof(uint32_t) * (2 * FCZ_NTIER + 2), e->stream));
    return FCZ_OK;
}

static int run_scan(fcz_engine* e, ScanArgs& sa) {
    uint32_t ntiles = (sa.n + SCAN_TILE - 1) / SCAN_TILE;
    if (ntiles == 0) ntiles = 1;
    sa.partial = (uint64_t*)e->partial.p;
    sa.totals = e->d_totals;
    k_scan_partials<<<ntiles, SCAN_THREADS, 0, e->stream>>>(sa);
    k_scan_final<<<ntiles, SCAN_THREADS, 0, e->stream>>>(sa);
    e->launches += 2;
    CK(cudaGetLastError());
    return FCZ_OK;
}

// counters + totals -> pinned host, then wait for them
static int fetch_plan(fcz_engine* e) {
    CK(cudaMemcpyAsync(e->h_counters, e->d_counters, sizeof(uint32_t) * (2 * FCZ_NTIER + 2), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(e->h_totals, e->d_totals, sizeof(uint64_t) * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return FCZ_OK;
}

// ------------------------------------------------------------------------------------------- encode

// one tier's persistent launch: k_encode, or k_encode_long with its global workspace for the long tier
static int launch_encode(fcz_engine* e, EncArgs& a, int tier, uint32_t cnt, uint32_t long_max_res, cudaStream_t st) {
    a.cfg = e->enc_tier[tier];
    a.gws = nullptr; a.gws_stride = 0; a.gws_max_res = 0;
    a.prefetch_next = e->enc_prefetch_next;
    uint32_t grid = (uint32_t)(e->num_sms * e->enc_occ[tier]);
    if (grid > cnt) grid = cnt;
    ProfSpan pk(e, FCZ_PROF_K_ENCODE, st);
    if (a.cfg.gws) {
        if (grid > (uint32_t)e->num_sms) grid = (uint32_t)e->num_sms;
        if (long_max_res < 1u || long_max_res > a.cfg.max_res) long_max_res = a.cfg.max_res;
        a.gws_stride = enc_gws_bytes(long_max_res);
        a.gws_max_res = long_max_res;
        int rc;
        if ((rc = ensure(e, e->enc_gws, a.gws_stride * grid))) return rc;  // stream-ordered reuse: launches on one engine are serialised
        a.gws = (uint8_t*)e->enc_gws.p;
        k_encode_long<<<grid, a.cfg.threads, a.cfg.smem, st>>>(a);
    } else {
        k_encode<<<grid, a.cfg.threads, a.cfg.smem, st>>>(a);
    }
    e->launches++;
    return FCZ_OK;
}

static int encode_device(fcz_engine* e, const fcz_chain_batch* in, fcz_blob_batch* out, uint64_t* total_bytes) {
    const uint32_t n = in->n_chains;
    int rc;
    if ((rc = plan_buffers(e, n))) return rc;
    TierTable tt;
    for (int i = 0; i < FCZ_NTIER; i++) tt.t[i] = e->enc_tier[i];
    PlanOut po;
    po.v0 = (uint32_t*)e->v0.p; po.v1 = nullptr; po.v2 = nullptr;
    po.status = out->status ? out->status : (int32_t*)e->status.p;
    po.tier_count = e->d_counters;
    po.tier_list = (uint32_t*)e->tier_list.p;
    static const bool use_desc = [] { const char* v = getenv("FCZ_ENC_DESC"); return !v || atoi(v) != 0; }();  // (A/B knob)
    if (use_desc && (rc = ensure(e, e->d_enc_desc, sizeof(EncDesc) * (size_t)FCZ_NTIER * n + 64))) return rc;
    po.enc_desc = use_desc ? (EncDesc*)e->d_enc_desc.p : nullptr;
    if (n) {
        k_enc_plan<<<(n + 7) / 8, 256, 0, e->stream>>>(n, in->res_off, in->atom_off, in->title_off, in->res_type,
                                                       e->opts.anchor_threshold, e->opts.terminate_blobs ? 1u : 0u, e->d_tables, tt, po);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 1;
    sa.in[0] = po.v0; sa.out[0] = out->blob_off; sa.out64[0] = 1;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = fetch_plan(e))) return rc;
    *total_bytes = e->h_totals[0];
    if (*total_bytes > out->bytes_cap) return fail(e, FCZ_E_CAPACITY, "encode needs %llu bytes, capacity %llu",
                                                   (unsigned long long)*total_bytes, (unsigned long long)out->bytes_cap);
    int ntier = 0;
    for (int i = 0; i < FCZ_NTIER; i++) ntier += e->h_counters[i] ? 1 : 0;
    // tiers run side by side on their own streams (the long-chain tiers have few, long-running blocks)
    const bool fork = ntier > 1;
    ProfSpan ps(e, FCZ_PROF_ENCODE);
    if (fork) CK(cudaEventRecord(e->ev_fork, e->stream));
    for (int i = 0; i < FCZ_NTIER; i++) {
        const uint32_t cnt = e->h_counters[i];
        if (!cnt) continue;
        cudaStream_t st = fork ? e->s_tier[i] : e->stream;
        if (fork) CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
        EncArgs a;
        a.res_off = in->res_off; a.atom_off = in->atom_off; a.title_off = in->title_off;
        a.res_type = in->res_type; a.bfactor = in->bfactor; a.xyz = in->xyz; a.titles = in->titles; a.meta = in->meta;
        a.blob_off = out->blob_off; a.bytes = out->bytes;
        a.list = (uint32_t*)e->tier_list.p + (size_t)i * n;
        a.desc = po.enc_desc ? po.enc_desc + (size_t)i * n : nullptr;
        a.count = e->d_counters + i;
        a.count_val = 0;
        a.ticket = e->d_counters + FCZ_NTIER + i;
        a.tables = e->d_tables;
        a.b = e->opts.anchor_threshold;
        a.term = e->opts.terminate_blobs ? 1u : 0u;
        if ((rc = launch_encode(e, a, i, cnt, e->h_counters[2 * FCZ_NTIER], st))) return rc;
        if (fork) {
            CK(cudaEventRecord(e->ev_join[i], st));
            CK(cudaStreamWaitEvent(e->stream, e->ev_join[i], 0));
        }
    }
    CK(cudaGetLastError());
    return FCZ_OK;
}


// ------------------------------------------------------------------- pipelined host-memory path
// Host batches are planned ON THE HOST (sizes, offsets, tiers: the offset arrays are host memory
// anyway) and processed in chunks so that the H2D copy of chunk k+1, the kernels of chunk k and the
// D2H copy of chunk k-1 overlap on three streams (PCIe is full duplex; the kernels hide behind it).

// Sum of table atom counts over L residue codes and whether any code has no table entry: the host-side twin of
// k_enc_plan's per-residue loop.  32 residues per step with AVX2 (two 16-entry byte shuffles as the lookup table)
// when the CPU has it; the planning of a host batch sits between its H2D copies and its kernels, so it is on the
// end-to-end critical path.
static void sum_natoms_scalar(const uint8_t* rt, uint32_t L, const uint8_t* lut, uint32_t* sum, uint32_t* bad) {
    uint32_t s = 0, b = 0;
    for (uint32_t r = 0; r < L; r++) { const uint32_t na = lut[rt[r]]; s += na; b |= (na == 0u); }
    *sum += s; *bad |= b;
}
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) static void sum_natoms_avx2(const uint8_t* rt, uint32_t L, const uint8_t* lut, uint32_t* sum, uint32_t* bad) {
    const __m128i lo128 = _mm_loadu_si128((const __m128i*)lut), hi128 = _mm_loadu_si128((const __m128i*)(lut + 16));
    const __m256i lut_lo = _mm256_broadcastsi128_si256(lo128), lut_hi = _mm256_broadcastsi128_si256(hi128);
    const __m256i m0f = _mm256_set1_epi8(0x0F), m10 = _mm256_set1_epi8(0x10), me0 = _mm256_set1_epi8((char)0xE0), zero = _mm256_setzero_si256();
    __m256i acc = zero, anybad = zero;
    uint32_t r = 0;
    for (; r + 32u <= L; r += 32u) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(rt + r));
        const __m256i idx = _mm256_and_si256(v, m0f);
        const __m256i is_hi = _mm256_cmpeq_epi8(_mm256_and_si256(v, m10), m10);
        __m256i na = _mm256_blendv_epi8(_mm256_shuffle_epi8(lut_lo, idx), _mm256_shuffle_epi8(lut_hi, idx), is_hi);
        const __m256i big = _mm256_cmpeq_epi8(_mm256_and_si256(v, me0), zero);  // 0xFF where the code is < 32
        na = _mm256_and_si256(na, big);
        anybad = _mm256_or_si256(anybad, _mm256_cmpeq_epi8(na, zero));
        acc = _mm256_add_epi64(acc, _mm256_sad_epu8(na, zero));
    }
    uint64_t t[4];
    _mm256_storeu_si256((__m256i*)t, acc);
    *sum += (uint32_t)(t[0] + t[1] + t[2] + t[3]);
    if (!_mm256_testz_si256(anybad, anybad)) *bad |= 1u;
    sum_natoms_scalar(rt + r, L - r, lut, sum, bad);
}
#endif
static void sum_natoms(const uint8_t* rt, uint32_t L, const uint8_t* lut, uint32_t* sum, uint32_t* bad) {
#if defined(__x86_64__) && defined(__GNUC__)
    static const bool has_avx2 = __builtin_cpu_supports("avx2");
    if (has_avx2) { sum_natoms_avx2(rt, L, lut, sum, bad); return; }
#endif
    sum_natoms_scalar(rt, L, lut, sum, bad);
}

static int host_pick_tier(const TierCfg* tiers, uint32_t L, uint64_t A, uint64_t blob, uint32_t nseg) {
    for (int i = 0; i < FCZ_NTIER; i++) {
        const TierCfg& t = tiers[i];
        if (L <= t.max_res && A <= t.max_atoms && blob <= t.max_blob && nseg <= t.max_seg) return i;
    }
    return -1;
}

static cudaEvent_t pool_event(fcz_engine* e, size_t i) {
    while (e->ev_pool.size() <= i) {
        cudaEvent_t ev = nullptr;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        e->ev_pool.push_back(ev);
    }
    return e->ev_pool[i];
}

// chain ranges of ~equal payload; chunk_c0 has nchunks+1 entries
static void make_chunks(uint32_t n, const uint64_t* weight_prefix /* [n+1] atoms */, std::vector<uint32_t>& chunk_c0, uint64_t chunk_bytes) {
    const uint64_t total = weight_prefix[n] - weight_prefix[0];
    uint32_t nchunks = (uint32_t)(total * 12ull / chunk_bytes) + 1u;  // coordinates (12 B per atom) per chunk
    if (nchunks > FCZ_HOST_CHUNKS) nchunks = FCZ_HOST_CHUNKS;
    if (nchunks > n) nchunks = n ? n : 1u;
    chunk_c0.assign(1, 0u);
    uint32_t c = 0;
    for (uint32_t k = 1; k < nchunks; k++) {
        const uint64_t target = weight_prefix[0] + total * k / nchunks;
        while (c < n && weight_prefix[c] < target) c++;
        if (c > chunk_c0.back()) chunk_c0.push_back(c);
    }
    chunk_c0.push_back(n);
}

// group chains by (chunk, tier): fills plan.list and plan.launches from per-chain tiers (-1 = skipped)
static void group_launches(fcz_engine::HostPlan& plan, const std::vector<int8_t>& tier) {
    plan.list.clear();
    plan.launches.clear();
    const uint32_t nchunks = (uint32_t)plan.chunk_c0.size() - 1u;
    for (uint32_t k = 0; k < nchunks; k++) {
        for (int t = 0; t < FCZ_NTIER; t++) {
            const uint32_t first = (uint32_t)plan.list.size();
            for (uint32_t c = plan.chunk_c0[k]; c < plan.chunk_c0[k + 1]; c++)
                if (tier[c] == t) plan.list.push_back(c);
            const uint32_t cnt = (uint32_t)plan.list.size() - first;
            if (cnt) plan.launches.push_back({k, (uint32_t)t, first, cnt});
        }
    }
}

#define COPY(dst, src, bytes, kind, st)                                                \
    do {                                                                               \
        if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, kind, st));                     \
    } while (0)

static int encode_host(fcz_engine* e, const fcz_chain_batch* in, fcz_blob_batch* out) {
    const uint32_t n = in->n_chains;
    int rc;
    const uint64_t n_res = in->res_off[n], n_atoms = in->atom_off[n], n_title = in->title_off[n];
    const int32_t b = e->opts.anchor_threshold;
    if ((rc = ensure(e, e->d_res_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_atom_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_title_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_res_type, n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_bfactor, 4ull * n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_xyz, 12ull * n_atoms + 16))) return rc;
    if ((rc = ensure(e, e->d_titles, n_title + 16))) return rc;
    if ((rc = ensure(e, e->d_meta, sizeof(fcz_chain_meta) * (uint64_t)n + 16))) return rc;
    if ((rc = ensure(e, e->d_blob_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_list, 4ull * n + 16))) return rc;
    if ((rc = ensure(e, e->d_tickets, 4ull * FCZ_HOST_CHUNKS * FCZ_NTIER))) return rc;

    fcz_engine::HostPlan plan;
    plan.n = n;
    make_chunks(n, in->atom_off, plan.chunk_c0, e->chunk_bytes);
    const uint32_t nchunks = (uint32_t)plan.chunk_c0.size() - 1u;

    // 1. start moving the inputs (copy stream), chunk by chunk; with a non-zero h2d_depth only that many chunks are
    // queued ahead of the kernels (the copy engine serves its queue in order, see the note at FCZ_CHUNK_BYTES).
    size_t evi = 0;
    cudaEvent_t ev0 = pool_event(e, evi++);
    CK(cudaEventRecord(ev0, e->stream));
    CK(cudaStreamWaitEvent(e->s_in, ev0, 0));
    CK(cudaStreamWaitEvent(e->s_out, ev0, 0));
    COPY(e->d_res_off.p, in->res_off, 4ull * (n + 1), cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_atom_off.p, in->atom_off, 8ull * (n + 1), cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_title_off.p, in->title_off, 4ull * (n + 1), cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_meta.p, in->meta, sizeof(fcz_chain_meta) * (uint64_t)n, cudaMemcpyHostToDevice, e->s_in);
    // The per-residue side arrays (5 B/residue against 94 B/residue of coordinates) go up whole, ahead of the chunks:
    // every copy has a fixed cost that grows when the other PCIe direction is busy (tools/pcie_matrix.py: 1 MB copies
    // reach 29 GB/s each way, 12 MB copies 44 GB/s), so a chunk is ONE copy -- its coordinates.
    COPY(e->d_res_type.p, in->res_type, n_res, cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_bfactor.p, in->bfactor, 4ull * n_res, cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_titles.p, in->titles, n_title, cudaMemcpyHostToDevice, e->s_in);
    std::vector<cudaEvent_t> ev_in(nchunks);
    for (uint32_t k = 0; k < nchunks; k++) ev_in[k] = pool_event(e, evi++);
    auto issue_chunk = [&](uint32_t k) -> int {
        const uint32_t c0 = plan.chunk_c0[k], c1 = plan.chunk_c0[k + 1];
        const uint64_t a0 = in->atom_off[c0], a1 = in->atom_off[c1];
        COPY((float*)e->d_xyz.p + 3ull * a0, in->xyz + 3ull * a0, 12ull * (a1 - a0), cudaMemcpyHostToDevice, e->s_in);
        CK(cudaEventRecord(ev_in[k], e->s_in));
        return FCZ_OK;
    };
    const uint32_t h2d_depth = e->h2d_depth ? e->h2d_depth : nchunks;  // 0: everything queued at once
    for (uint32_t k = 0; k < nchunks && k < h2d_depth; k++)
        if ((rc = issue_chunk(k))) return rc;

    // 2. meanwhile plan on the host: validate (what k_enc_plan does on the device), sizes, offsets, tiers
    uint8_t nat_lut[256];
    for (int i = 0; i < 256; i++) nat_lut[i] = i < FCZ_NUM_CODES ? FCZ_NATOMS[i] : 0;
    std::vector<int8_t> tier(n, -1);
    uint32_t long_max_res = 0;  // longest chain of the long tier: sizes its global workspace
    plan.status.assign(n, FCZ_OK);
    out->blob_off[0] = 0;
    for (uint32_t c = 0; c < n; c++) {
        const uint32_t r0 = in->res_off[c], L = in->res_off[c + 1] - r0;
        const uint64_t A = in->atom_off[c + 1] - in->atom_off[c];
        const uint32_t T = in->title_off[c + 1] - in->title_off[c];
        int st = FCZ_OK;
        uint64_t size = 0;
        uint32_t sum = 0, bad = 0;
        const uint8_t* rt = in->res_type + r0;
        sum_natoms(rt, L, nat_lut, &sum, &bad);
        if (L < 2u || L > 65535u || b < 1) st = FCZ_E_LIMIT;
        else if (bad) st = FCZ_E_RESIDUE;
        else if ((uint64_t)sum != A) st = FCZ_E_ARG;
        else {
            const int na = anchor_count(L, b);
            if (na > 255) st = FCZ_E_LIMIT;
            else {
                size = make_layout(L, sum - 3u * L, T, (uint32_t)na).size + (e->opts.terminate_blobs ? 1u : 0u);
                const int t = host_pick_tier(e->enc_tier, L, sum, size, (uint32_t)na - 1u);
                if (t < 0) { st = FCZ_E_LIMIT; size = 0; }
                else {
                    tier[c] = (int8_t)t;
                    if (e->enc_tier[t].gws && L > long_max_res) long_max_res = L;
                }
            }
        }
        plan.status[c] = st;
        if (out->status) out->status[c] = st;
        out->blob_off[c + 1] = out->blob_off[c] + size;
    }
    const uint64_t total = out->blob_off[n];
    if (total > out->bytes_cap) {
        cudaStreamSynchronize(e->s_in);
        return fail(e, FCZ_E_CAPACITY, "encode needs %llu bytes, capacity %llu", (unsigned long long)total,
                    (unsigned long long)out->bytes_cap);
    }
    if ((rc = ensure(e, e->d_bytes, total + 64))) return rc;
    group_launches(plan, tier);
    COPY(e->d_blob_off.p, out->blob_off, 8ull * (n + 1), cudaMemcpyHostToDevice, e->s_in);
    COPY(e->d_list.p, plan.list.data(), 4ull * plan.list.size(), cudaMemcpyHostToDevice, e->s_in);
    cudaEvent_t ev_plan = pool_event(e, evi++);
    CK(cudaEventRecord(ev_plan, e->s_in));
    CK(cudaMemsetAsync(e->d_tickets.p, 0, 4ull * FCZ_HOST_CHUNKS * FCZ_NTIER, e->stream));
    CK(cudaStreamWaitEvent(e->stream, ev_plan, 0));

    // 3. kernels per chunk (main stream) and the blobs back (out stream)
    size_t li = 0;
    for (uint32_t k = 0; k < nchunks; k++) {
        if (k + h2d_depth < nchunks) {  // chunk k has landed: queue the next one behind the one in flight
            CK(cudaEventSynchronize(ev_in[k]));
            if ((rc = issue_chunk(k + h2d_depth))) return rc;
        }
        CK(cudaStreamWaitEvent(e->stream, ev_in[k], 0));
        for (; li < plan.launches.size() && plan.launches[li].chunk == k; li++) {
            const fcz_engine::Launch& ln = plan.launches[li];
            EncArgs a;
            a.res_off = (uint32_t*)e->d_res_off.p; a.atom_off = (uint64_t*)e->d_atom_off.p; a.title_off = (uint32_t*)e->d_title_off.p;
            a.res_type = (uint8_t*)e->d_res_type.p; a.bfactor = (float*)e->d_bfactor.p; a.xyz = (float*)e->d_xyz.p;
            a.titles = (char*)e->d_titles.p; a.meta = (fcz_chain_meta*)e->d_meta.p;
            a.blob_off = (uint64_t*)e->d_blob_off.p; a.bytes = (uint8_t*)e->d_bytes.p;
            a.list = (uint32_t*)e->d_list.p + ln.first;
            a.desc = nullptr;  // host-planned: the kernel reads the offset arrays
            a.count = nullptr; a.count_val = ln.count;
            a.ticket = (uint32_t*)e->d_tickets.p + k * FCZ_NTIER + ln.tier;
            a.tables = e->d_tables; a.b = b; a.term = e->opts.terminate_blobs ? 1u : 0u;
            {
                ProfSpan ps(e, FCZ_PROF_ENCODE);
                if ((rc = launch_encode(e, a, (int)ln.tier, ln.count, long_max_res, e->stream))) return rc;
            }
        }
        if ((k + 1u) % FCZ_D2H_GROUP == 0u || k + 1u == nchunks) {  // the blobs of the last few chunks, one copy
            const uint32_t k0 = k - k % FCZ_D2H_GROUP;
            cudaEvent_t ev_k = pool_event(e, evi++);
            CK(cudaEventRecord(ev_k, e->stream));
            CK(cudaStreamWaitEvent(e->s_out, ev_k, 0));
            const uint64_t b0 = out->blob_off[plan.chunk_c0[k0]], b1 = out->blob_off[plan.chunk_c0[k + 1]];
            COPY(out->bytes + b0, (uint8_t*)e->d_bytes.p + b0, b1 - b0, cudaMemcpyDeviceToHost, e->s_out);
        }
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->s_out));
    CK(cudaStreamSynchronize(e->stream));
    return FCZ_OK;
}

#define H2D(buf, src, bytes)                                                                   \
    do {                                                                                       \
        if ((rc = ensure(e, buf, (bytes) + 16))) return rc;                                    \
        if (bytes) CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, e->stream)); \
    } while (0)

extern "C" int fcz_encode_batch(fcz_engine* e, const fcz_chain_batch* in, fcz_blob_batch* out) {
    if (!e || !in || !out) return FCZ_E_ARG;
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    out->n_chains = in->n_chains;
    const uint32_t n = in->n_chains;
    uint64_t total = 0;
    if (in->mem == FCZ_MEM_DEVICE) return encode_device(e, in, out, &total);
    return encode_host(e, in, out);
}

// --------------------------------------------------------------------------------------------- text

// arrays of a chain batch on the device: the batch itself (device memory) or the engine's staging copy of a host batch
struct DevChains {
    const uint32_t* res_off; const uint64_t* atom_off; const uint32_t* title_off;
    const uint8_t* res_type; const float* bfactor; const float* xyz; const char* titles; const fcz_chain_meta* meta;
};

static int upload_chains(fcz_engine* e, const fcz_chain_batch* in, DevChains* d) {
    const uint32_t n = in->n_chains;
    if (in->mem == FCZ_MEM_DEVICE) {
        d->res_off = in->res_off; d->atom_off = in->atom_off; d->title_off = in->title_off; d->res_type = in->res_type;
        d->bfactor = in->bfactor; d->xyz = in->xyz; d->titles = in->titles; d->meta = in->meta;
        return FCZ_OK;
    }
    int rc;
    const uint64_t n_res = in->res_off[n], n_atoms = in->atom_off[n], n_title = in->title_off[n];
    H2D(e->d_res_off, in->res_off, 4ull * (n + 1));
    H2D(e->d_atom_off, in->atom_off, 8ull * (n + 1));
    H2D(e->d_title_off, in->title_off, 4ull * (n + 1));
    H2D(e->d_res_type, in->res_type, n_res);
    H2D(e->d_bfactor, in->bfactor, 4ull * n_res);
    H2D(e->d_xyz, in->xyz, 12ull * n_atoms);
    H2D(e->d_titles, in->titles, n_title);
    H2D(e->d_meta, in->meta, sizeof(fcz_chain_meta) * (uint64_t)n);
    d->res_off = (uint32_t*)e->d_res_off.p; d->atom_off = (uint64_t*)e->d_atom_off.p; d->title_off = (uint32_t*)e->d_title_off.p;
    d->res_type = (uint8_t*)e->d_res_type.p; d->bfactor = (float*)e->d_bfactor.p; d->xyz = (float*)e->d_xyz.p;
    d->titles = (char*)e->d_titles.p; d->meta = (fcz_chain_meta*)e->d_meta.p;
    return FCZ_OK;
}

static void pdb_args(fcz_engine* e, const DevChains& d, uint32_t n, PdbArgs* a) {
    memset(a, 0, sizeof *a);
    a->res_off = d.res_off; a->atom_off = d.atom_off; a->title_off = d.title_off; a->res_type = d.res_type; a->bfactor = d.bfactor;
    a->xyz = d.xyz; a->titles = d.titles; a->meta = d.meta;
    a->tt = e->d_text_tables; a->use_alt = e->opts.use_alt_atom_order; a->n = n;
    a->aoff = (uint32_t*)e->ws_aoff.p; a->toff = (uint32_t*)e->ws_toff.p;
    a->text_bytes = (uint32_t*)e->v0.p; a->units = (uint32_t*)e->v1.p;
    a->unit_off = (uint32_t*)e->d_unit_off.p;
    a->unit_chain = (uint32_t*)e->d_unit_chain.p;
}

// plan over device-resident chains: fills d_text_off (device, [n+1]) and the engine's unit offsets; syncs; leaves the
// totals in e->pdb.  host_text_off (may be null) receives a copy of the offsets, and the unit offsets come to the host too.
static int pdb_plan_dev(fcz_engine* e, const DevChains& d, uint32_t n, uint64_t n_res_cap, uint64_t* d_text_off, uint64_t* host_text_off) {
    int rc;
    e->pdb.valid = false;
    if ((rc = plan_buffers(e, n))) return rc;
    if ((rc = ensure(e, e->ws_aoff, 4ull * (n_res_cap + n + 1)))) return rc;
    if ((rc = ensure(e, e->ws_toff, 4ull * (n_res_cap + n + 1)))) return rc;
    if ((rc = ensure(e, e->d_unit_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_unit_chain, 4ull * (n_res_cap / FCZ_PDB_UNIT_RES + n + 1)))) return rc;  // units <= residues/UNIT + chains
    PdbArgs a;
    pdb_args(e, d, n, &a);
    if (n) {
        ProfSpan pk(e, FCZ_PROF_K_PDB_PLAN);
        k_pdb_plan<<<n, 128, 0, e->stream>>>(a);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 2;
    sa.in[0] = a.text_bytes; sa.out[0] = d_text_off; sa.out64[0] = 1;
    sa.in[1] = a.units; sa.out[1] = e->d_unit_off.p; sa.out64[1] = 0;
    if ((rc = run_scan(e, sa))) return rc;
    if (n) {
        k_pdb_unit_map<<<(n + 255) / 256, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    if (host_text_off) {
        e->h_unit_off.resize((size_t)n + 1);
        CK(cudaMemcpyAsync(host_text_off, d_text_off, 8ull * (n + 1), cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(e->h_unit_off.data(), e->d_unit_off.p, 4ull * (n + 1), cudaMemcpyDeviceToHost, e->stream));
    }
    if ((rc = fetch_plan(e))) return rc;
    e->pdb.n = n; e->pdb.total_bytes = e->h_totals[0]; e->pdb.total_units = (uint32_t)e->h_totals[1]; e->pdb.valid = true;
    return FCZ_OK;
}

// emit into the engine's text buffer in slabs of ~64 MB so that the D2H copy of slab k overlaps the kernel of slab k+1
static int pdb_emit_to_host(fcz_engine* e, const DevChains& d, uint32_t n, const uint64_t* host_text_off, char* host_bytes) {
    int rc;
    if ((rc = ensure(e, e->d_text, e->pdb.total_bytes + 64))) return rc;
    PdbArgs a;
    pdb_args(e, d, n, &a);
    a.text_off = (uint64_t*)e->d_text_off.p; a.text = (char*)e->d_text.p;
    size_t evi = 0;
    cudaEvent_t ev0 = pool_event(e, evi++);
    CK(cudaEventRecord(ev0, e->stream));
    CK(cudaStreamWaitEvent(e->s_out, ev0, 0));
    uint32_t c0 = 0;
    while (c0 < n) {
        uint32_t c1 = c0;
        while (c1 < n && host_text_off[c1] - host_text_off[c0] < (64ull << 20)) c1++;
        const uint32_t u0 = e->h_unit_off[c0], u1 = e->h_unit_off[c1];
        if (u1 > u0) {
            a.unit0 = u0;
            ProfSpan pk(e, FCZ_PROF_K_PDB_EMIT);
            k_pdb_emit<<<u1 - u0, 128, 0, e->stream>>>(a);
            e->launches++;
        }
        cudaEvent_t ev = pool_event(e, evi++);
        CK(cudaEventRecord(ev, e->stream));
        CK(cudaStreamWaitEvent(e->s_out, ev, 0));
        const uint64_t b0 = host_text_off[c0], b1 = host_text_off[c1];
        COPY(host_bytes + b0, (char*)e->d_text.p + b0, b1 - b0, cudaMemcpyDeviceToHost, e->s_out);
        c0 = c1;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->s_out));
    CK(cudaStreamSynchronize(e->stream));
    return FCZ_OK;
}

static DevChains staged_chains(fcz_engine* e) {
    DevChains d;
    d.res_off = (uint32_t*)e->d_res_off.p; d.atom_off = (uint64_t*)e->d_atom_off.p; d.title_off = (uint32_t*)e->d_title_off.p;
    d.res_type = (uint8_t*)e->d_res_type.p; d.bfactor = (float*)e->d_bfactor.p; d.xyz = (float*)e->d_xyz.p;
    d.titles = (char*)e->d_titles.p; d.meta = (fcz_chain_meta*)e->d_meta.p;
    return d;
}

extern "C" int fcz_pdb_text_plan(fcz_engine* e, const fcz_chain_batch* in, fcz_text_batch* out, uint64_t* total_bytes) {
    if (!e || !in || !out || !total_bytes) return FCZ_E_ARG;
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    e->pdb.valid = false;
    int rc;
    DevChains d;
    if ((rc = upload_chains(e, in, &d))) return rc;
    const bool host = in->mem == FCZ_MEM_HOST;
    if (host && (rc = ensure(e, e->d_text_off, 8ull * (n + 1)))) return rc;
    if ((rc = pdb_plan_dev(e, d, n, host ? in->res_off[n] : in->res_cap, host ? (uint64_t*)e->d_text_off.p : out->text_off,
                           host ? out->text_off : nullptr))) return rc;
    *total_bytes = e->pdb.total_bytes;
    return FCZ_OK;
}

extern "C" int fcz_pdb_text_batch(fcz_engine* e, const fcz_chain_batch* in, fcz_text_batch* out) {
    if (!e || !in || !out) return FCZ_E_ARG;
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    if (!e->pdb.valid || e->pdb.n != n) return fail(e, FCZ_E_ARG, "fcz_pdb_text_batch needs a preceding fcz_pdb_text_plan on the same batch");
    if (e->pdb.total_bytes > out->bytes_cap)
        return fail(e, FCZ_E_CAPACITY, "text needs %llu bytes, capacity %llu", (unsigned long long)e->pdb.total_bytes, (unsigned long long)out->bytes_cap);
    e->pdb.valid = false;
    int rc;
    if (in->mem == FCZ_MEM_HOST) return pdb_emit_to_host(e, staged_chains(e), n, out->text_off, out->bytes);  // chains still staged since the plan
    DevChains d;
    if ((rc = upload_chains(e, in, &d))) return rc;
    PdbArgs a;
    pdb_args(e, d, n, &a);
    a.text_off = out->text_off; a.text = out->bytes; a.unit0 = 0;
    if (e->pdb.total_units) {
        ProfSpan pk(e, FCZ_PROF_K_PDB_EMIT);
        k_pdb_emit<<<e->pdb.total_units, 128, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    return FCZ_OK;
}

extern "C" int fcz_extract_batch(fcz_engine* e, const fcz_blob_batch* in, int32_t type, int32_t digits, fcz_text_batch* out, uint64_t* total_bytes) {
    if (!e || !in || !out || !total_bytes) return FCZ_E_ARG;
    if (type != 0 && type != 1) return fail(e, FCZ_E_ARG, "extract type must be 0 (plddt) or 1 (sequence)");
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    int rc;
    if (digits < 1) digits = 1; else if (digits > 4) digits = 4;  // src/foldcomp.cpp:1265-1269
    if ((rc = plan_buffers(e, n))) return rc;
    ExtractArgs a;
    memset(&a, 0, sizeof a);
    const bool host = in->mem == FCZ_MEM_HOST;
    if (host) {
        const uint64_t nb = in->blob_off[n];
        H2D(e->d_blob_off, in->blob_off, 8ull * (n + 1));
        H2D(e->d_bytes, in->bytes, nb);
        if ((rc = ensure(e, e->d_text_off, 8ull * (n + 1)))) return rc;
        a.blob_off = (uint64_t*)e->d_blob_off.p; a.bytes = (uint8_t*)e->d_bytes.p;
    } else {
        a.blob_off = in->blob_off; a.bytes = in->bytes;
    }
    a.tt = e->d_text_tables; a.type = type; a.digits = (uint32_t)digits; a.n = n;
    a.text_bytes = (uint32_t*)e->v0.p; a.status = (int32_t*)e->status.p;
    uint64_t* d_text_off = host ? (uint64_t*)e->d_text_off.p : out->text_off;
    if (n) {
        k_extract_plan<<<(n + 255) / 256, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 1;
    sa.in[0] = a.text_bytes; sa.out[0] = d_text_off; sa.out64[0] = 1;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = fetch_plan(e))) return rc;
    const uint64_t total = e->h_totals[0];
    *total_bytes = total;
    if (total > out->bytes_cap) return fail(e, FCZ_E_CAPACITY, "extract needs %llu bytes, capacity %llu", (unsigned long long)total, (unsigned long long)out->bytes_cap);
    if (host && (rc = ensure(e, e->d_text, total + 64))) return rc;
    a.text_off = d_text_off; a.text = host ? (char*)e->d_text.p : out->bytes;
    if (n) {
        k_extract<<<n, 128, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    if (host) {
        CK(cudaMemcpyAsync(out->text_off, d_text_off, 8ull * (n + 1), cudaMemcpyDeviceToHost, e->stream));
        if (total) CK(cudaMemcpyAsync(out->bytes, e->d_text.p, total, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return FCZ_OK;
}

// ------------------------------------------------------------------------------------------- decode

struct Dec2Tier {
    uint32_t count, max_L, max_anchor, max_blob, max_atoms;  // chains and maxima over them
    size_t list_off;                                         // first entry in the device chain list
    bool smem;                                               // runs the shared-memory kernels
};
struct Dec2Sub {
    uint32_t c0, c1, r0, r1, s0, s1;  // chains, residues, segment slots of one sub-batch
    Dec2Tier tier[FCZ_DEC_TIERS];
};

#define FCZ_SMEM_LIMIT (227u * 1024u)

static uint32_t stitch_chain_bytes(uint32_t max_anchor) { return 4u * ((max_anchor * (uint32_t)SegPacked::N) | 1u); }

static int dec2_workspace(fcz_engine* e, Dec2Sub* subs, size_t nsub) {
    uint64_t mr = 0, ms = 0, mc = 0;
    bool any_global = false, any_smem = false;
    const bool force_global = getenv("FCZ_DEC_GLOBAL") != nullptr;
    for (size_t k = 0; k < nsub; k++) {
        const uint64_t nr = subs[k].r1 - subs[k].r0, nc = subs[k].c1 - subs[k].c0;
        if (nc > mc) mc = nc;
        if (nr > mr) mr = nr;
        if ((uint64_t)(subs[k].s1 - subs[k].s0) > ms) ms = subs[k].s1 - subs[k].s0;
        for (int t = 0; t < FCZ_DEC_TIERS; t++) {
            Dec2Tier& tr = subs[k].tier[t];
            if (!tr.count) continue;
            const FrontSmem fs = front_smem(tr.max_L, tr.max_anchor, tr.max_blob);
            const BackSmem bs = back_smem(tr.max_L, tr.max_anchor, tr.max_atoms);
            tr.smem = fs.total <= FCZ_SMEM_LIMIT && bs.total <= FCZ_SMEM_LIMIT && !force_global;
            (tr.smem ? any_smem : any_global) = true;
        }
    }
    int rc;
    if ((rc = ensure(e, e->sc_aoff, 4ull * (mr + mc + 1)))) return rc;
    if ((rc = ensure(e, e->sc_rev, 36ull * mr + 16))) return rc;
    if ((rc = ensure(e, e->sc_seg, 4ull * FCZ_SEG_FLOATS * (ms + 1)))) return rc;
    if (any_smem && (rc = ensure(e, e->sc_loc, 36ull * mr + 16))) return rc;
    if (any_global) {
        if ((rc = ensure(e, e->sc_segid, mr + 16))) return rc;
        if ((rc = ensure(e, e->sc_tor, 24ull * mr + 16))) return rc;
        if ((rc = ensure(e, e->sc_ang, 24ull * mr + 16))) return rc;
    }
    return FCZ_OK;
}

// the phase kernels over one sub-batch (a carries the batch pointers, list = device chain list)
static int dec2_launch(fcz_engine* e, Dec2Args a, const Dec2Sub& sb, const uint32_t* list, const DecDesc* desc = nullptr) {
    const uint32_t nch = sb.c1 - sb.c0;
    if (!nch) return FCZ_OK;
    a.c0 = sb.c0; a.c1 = sb.c1; a.r_base = sb.r0; a.s_base = sb.s0;
    a.aoff = (uint32_t*)e->sc_aoff.p; a.rev = (float*)e->sc_rev.p; a.seg = (float*)e->sc_seg.p;
    a.loc = (float*)e->sc_loc.p;
    a.segid = (uint8_t*)e->sc_segid.p; a.tor = (cs*)e->sc_tor.p; a.ang = (cs*)e->sc_ang.p;
    int ntier = 0;
    for (int t = 0; t < FCZ_DEC_TIERS; t++)
        if (sb.tier[t].count) ntier++;
    if (!ntier) return FCZ_OK;  // no valid chain
    // Each tier is its own three-kernel pipeline; tiers run side by side on their own streams (a tier of long
    // chains has few, long-running blocks): fork from the engine's stream, join at the end.  A tier with many chains
    // runs as TWO such pipelines over the halves of its list (streams t and 4 + t): the serial stitch of one half -- one
    // wave of latency-bound threads -- then overlaps the front / back kernels of the other half, and the tail of one
    // kernel fills with the head of the next.  The workspace is indexed per chain, so the halves share nothing.
    // Measured on the B200 (profiles/r02_v15_ab.jsonl): decode span 0.457 ms split against 0.463 ms whole on the headline
    // batch, 0.831 against 0.824 ms on mixed lengths -- no gain worth the overlapped per-kernel timings, so it is OFF
    // unless FCZ_DEC_SPLIT_MIN names a chain count.
    static const uint32_t split_min = [] { const char* v = getenv("FCZ_DEC_SPLIT_MIN"); return v ? (uint32_t)atol(v) : 0xFFFFFFFFu; }();
    bool fork = ntier > 1;
    for (int t = 0; t < FCZ_DEC_TIERS; t++) fork |= sb.tier[t].count >= split_min && sb.tier[t].smem;
    if (fork) CK(cudaEventRecord(e->ev_fork, e->stream));
    for (int t = 0; t < FCZ_DEC_TIERS; t++) {
        const Dec2Tier& tr = sb.tier[t];
        if (!tr.count) continue;
        const int halves = (tr.smem && tr.count >= split_min) ? 2 : 1;
        a.max_L = tr.max_L; a.max_anchor = tr.max_anchor; a.max_blob = tr.max_blob; a.max_atoms = tr.max_atoms;
        for (int h = 0; h < halves; h++) {
            cudaStream_t st = fork ? e->s_tier[t + 4 * h] : e->stream;
            if (fork) CK(cudaStreamWaitEvent(st, e->ev_fork, 0));
            const uint32_t first = h ? tr.count / 2u : 0u, cnt = halves == 2 ? (h ? tr.count - tr.count / 2u : tr.count / 2u) : tr.count;
            a.list = list + tr.list_off + first; a.count = cnt;
            a.desc = desc ? desc + tr.list_off + first : nullptr;
            // stitch: one wave when it fits -- chains per block = ceil(chains / SMs), bounded by shared memory
            const uint32_t cbytes = stitch_chain_bytes(tr.max_anchor);
            uint32_t G = (cnt + (uint32_t)e->num_sms - 1u) / (uint32_t)e->num_sms;
            const uint32_t gmax = (FCZ_SMEM_LIMIT - 1024u) / cbytes;
            if (G > gmax) G = gmax;
            if (G > 1024u) G = 1024u;
            if (G < 1u) G = 1u;
            a.stitch_group = G;
            uint32_t sthr = 32u * G;  // enough warps to gather the group's scratch with deep queues
            sthr = sthr < 256u ? 256u : (sthr > 1024u ? 1024u : sthr);
            if (tr.smem) {
                // front: three lanes per segment and direction -- enough warps that the passes take one trip
                uint32_t thr = (6u * tr.max_anchor + 31u) & ~31u;
                thr = thr < 128u ? 128u : (thr > 1024u ? 1024u : thr);
                {
                    ProfSpan pk(e, FCZ_PROF_K_DEC_FRONT, st);
                    k_dec_front<<<cnt, thr, front_smem(tr.max_L, tr.max_anchor, tr.max_blob).total, st>>>(a);
                }
                {
                    ProfSpan pk(e, FCZ_PROF_K_DEC_STITCH, st);
                    k_dec_stitch_t<<<(cnt + G - 1u) / G, sthr, G * cbytes, st>>>(a);
                }
                // back: about one thread per two residues (side chains take two trips), whole warps
                thr = ((tr.max_L * 35u) / 64u + 31u) & ~31u;
                thr = thr < 192u ? 192u : (thr > 1024u ? 1024u : thr);
                {
                    ProfSpan pk(e, FCZ_PROF_K_DEC_BACK, st);
                    k_dec_back<<<cnt, thr, back_smem(tr.max_L, tr.max_anchor, tr.max_atoms).total, st>>>(a);
                }
                e->launches += 3;
            } else {
                k_dec_unpack<<<cnt, 256, 0, st>>>(a);
                k_dec_passes<<<cnt, 128, 0, st>>>(a);
                k_dec_stitch_t<<<(cnt + G - 1u) / G, sthr, G * cbytes, st>>>(a);
                k_dec_blend<<<cnt, 128, 0, st>>>(a);
                k_dec_side<<<cnt, 192, 0, st>>>(a);
                e->launches += 5;
            }
            if (fork) {
                CK(cudaEventRecord(e->ev_join[t + 4 * h], st));
                CK(cudaStreamWaitEvent(e->stream, e->ev_join[t + 4 * h], 0));
            }
        }
    }
    return FCZ_OK;
}

// Blob headers are parsed on the host (Foldcomp::read, src/foldcomp.cpp:904-924): sizes and offsets.
// The per-residue checks (codes, anchors, side-chain count) run on the device before the decode kernels.
static int decode_plan_host(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals) {
    const uint32_t n = in->n_chains;
    fcz_engine::HostPlan& plan = e->hplan;
    plan.n = n;
    plan.status.assign(n, FCZ_OK);
    plan.seg_off.assign(n + 1, 0u);
    out->res_off[0] = 0; out->atom_off[0] = 0; out->title_off[0] = 0;
    for (uint32_t c = 0; c < n; c++) {
        const uint8_t* blob = in->bytes + in->blob_off[c];
        const uint64_t len = in->blob_off[c + 1] - in->blob_off[c];
        int st = FCZ_OK;
        uint32_t L = 0, T = 0, na = 0;
        uint64_t A = 0;
        if (len < HDR_BYTES || memcmp(blob, "FCMP", 4) != 0) st = FCZ_E_MAGIC;
        else {
            L = get_u16(blob + OFF_NRES);
            T = get_u32(blob + OFF_LENTITLE);
            const uint32_t nsc = get_u32(blob + OFF_NSC);
            na = blob[OFF_NANCHOR];
            const Layout y = make_layout(L, nsc, T, na);
            if (L < 2u || na < 2u || (uint64_t)T > len || (uint64_t)nsc > len || (uint64_t)y.size > len) st = FCZ_E_TRUNCATED;
            else A = (uint64_t)nsc + 3ull * L;  // = sum of table atoms when the blob is consistent (checked on the device)
        }
        if (st != FCZ_OK) { L = 0; A = 0; T = 0; na = 0; }
        plan.status[c] = st;
        if (out->status) out->status[c] = st;
        out->res_off[c + 1] = out->res_off[c] + L;
        out->atom_off[c + 1] = out->atom_off[c] + A;
        out->title_off[c + 1] = out->title_off[c] + T;
        plan.seg_off[c + 1] = plan.seg_off[c] + na;
    }
    make_chunks(n, out->atom_off, plan.chunk_c0, e->chunk_bytes);
    totals->n_res = out->res_off[n];
    totals->n_atoms = out->atom_off[n];
    totals->n_title_bytes = out->title_off[n];
    totals->n_blob_bytes = in->blob_off[n];
    return FCZ_OK;
}

static int decode_host(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out) {
    const uint32_t n = in->n_chains;
    fcz_engine::HostPlan& plan = e->hplan;
    if (plan.n != n || plan.chunk_c0.empty() || plan.seg_off.size() != (size_t)n + 1)
        return fail(e, FCZ_E_ARG, "fcz_decode_batch(host) needs a preceding fcz_decode_plan on the same batch");
    int rc;
    const uint64_t n_res = out->res_off[n], n_atoms = out->atom_off[n], n_title = out->title_off[n], n_bytes = in->blob_off[n];
    if (n_res > out->res_cap || n_atoms > out->atom_cap || (out->titles && n_title > out->title_cap))
        return fail(e, FCZ_E_CAPACITY, "decode output capacity too small");
    if ((rc = ensure(e, e->d_blob_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_bytes, n_bytes + 64))) return rc;
    if ((rc = ensure(e, e->d_res_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_atom_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_title_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_seg_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_status, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->d_res_type, n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_bfactor, 4ull * n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_xyz, 12ull * n_atoms + 16))) return rc;
    if ((rc = ensure(e, e->d_titles, n_title + 16))) return rc;
    if ((rc = ensure(e, e->d_meta, sizeof(fcz_chain_meta) * (uint64_t)n + 16))) return rc;
    const uint32_t nchunks = (uint32_t)plan.chunk_c0.size() - 1u;
    std::vector<Dec2Sub> subs(nchunks);
    plan.list.assign(n, 0u);  // chains grouped by (chunk, length tier)
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t c0 = plan.chunk_c0[k], c1 = plan.chunk_c0[k + 1];
        Dec2Sub& sb = subs[k];
        memset(&sb, 0, sizeof sb);
        sb.c0 = c0; sb.c1 = c1; sb.r0 = out->res_off[c0]; sb.r1 = out->res_off[c1]; sb.s0 = plan.seg_off[c0]; sb.s1 = plan.seg_off[c1];
        for (uint32_t c = c0; c < c1; c++) {
            if (plan.status[c] != FCZ_OK) continue;
            Dec2Tier& tr = sb.tier[dec_tier_of(out->res_off[c + 1] - out->res_off[c])];
            tr.count++;
            tr.max_L = std::max(tr.max_L, out->res_off[c + 1] - out->res_off[c]);
            tr.max_anchor = std::max(tr.max_anchor, plan.seg_off[c + 1] - plan.seg_off[c]);
            tr.max_blob = std::max(tr.max_blob, (uint32_t)(in->blob_off[c + 1] - in->blob_off[c]));
            tr.max_atoms = std::max(tr.max_atoms, (uint32_t)(out->atom_off[c + 1] - out->atom_off[c]));
        }
        size_t off = c0;
        uint32_t cur[FCZ_DEC_TIERS];
        for (int t = 0; t < FCZ_DEC_TIERS; t++) { sb.tier[t].list_off = off; cur[t] = 0; off += sb.tier[t].count; }
        for (uint32_t c = c0; c < c1; c++) {
            if (plan.status[c] != FCZ_OK) continue;
            const uint32_t t = dec_tier_of(out->res_off[c + 1] - out->res_off[c]);
            plan.list[sb.tier[t].list_off + cur[t]++] = c;
        }
    }
    if ((rc = ensure(e, e->d_dec_list, 4ull * n + 16))) return rc;
    if ((rc = dec2_workspace(e, subs.data(), subs.size()))) return rc;

    size_t evi = 0;
    cudaEvent_t ev0 = pool_event(e, evi++);
    CK(cudaEventRecord(ev0, e->stream));
    CK(cudaStreamWaitEvent(e->s_in, ev0, 0));
    CK(cudaStreamWaitEvent(e->s_out, ev0, 0));
    // Inputs go up in as FEW copies as possible: the seven per-chain arrays are packed into one pinned staging buffer
    // (one copy), the blobs follow in FCZ_DEC_H2D_PIECES copies: every copy has a fixed cost that grows when the other
    // PCIe direction is busy (tools/pcie_matrix.py).
    uint8_t* dst_stage = nullptr;
    {
        const uint64_t n1 = (uint64_t)n + 1;
        uint64_t off[8];
        uint64_t o = 0;
        const uint64_t sizes[7] = {8 * n1, 8 * n1, 4 * n1, 4 * n1, 4 * n1, 4ull * n, 4ull * n};
        for (int i = 0; i < 7; i++) { off[i] = o; o += (sizes[i] + 15u) & ~15ull; }
        off[7] = o;
        if (o > e->h_stage_cap) {
            if (e->h_stage) CK(cudaFreeHost(e->h_stage));
            e->h_stage = nullptr; e->h_stage_cap = 0;
            CK(cudaMallocHost(&e->h_stage, o + o / 4 + 4096));
            e->h_stage_cap = o + o / 4 + 4096;
        }
        if ((rc = ensure(e, e->d_stage, o + 16))) return rc;
        uint8_t* hs = (uint8_t*)e->h_stage;
        memcpy(hs + off[0], in->blob_off, sizes[0]);
        memcpy(hs + off[1], out->atom_off, sizes[1]);
        memcpy(hs + off[2], out->res_off, sizes[2]);
        memcpy(hs + off[3], out->title_off, sizes[3]);
        memcpy(hs + off[4], plan.seg_off.data(), sizes[4]);
        if (n) memcpy(hs + off[5], plan.status.data(), sizes[5]);
        if (n) memcpy(hs + off[6], plan.list.data(), sizes[6]);
        COPY(e->d_stage.p, hs, o, cudaMemcpyHostToDevice, e->s_in);
        dst_stage = (uint8_t*)e->d_stage.p;
        e->dh.blob_off = (uint64_t*)(dst_stage + off[0]); e->dh.atom_off = (uint64_t*)(dst_stage + off[1]);
        e->dh.res_off = (uint32_t*)(dst_stage + off[2]); e->dh.title_off = (uint32_t*)(dst_stage + off[3]);
        e->dh.seg_off = (uint32_t*)(dst_stage + off[4]); e->dh.status = (int32_t*)(dst_stage + off[5]);
        e->dh.list = (uint32_t*)(dst_stage + off[6]);
    }
    std::vector<cudaEvent_t> ev_in(nchunks);
    {
        const uint32_t npieces = nchunks < FCZ_DEC_H2D_PIECES ? nchunks : FCZ_DEC_H2D_PIECES;
        for (uint32_t p = 0; p < npieces; p++) {
            const uint32_t k0 = (uint32_t)((uint64_t)nchunks * p / npieces), k1 = (uint32_t)((uint64_t)nchunks * (p + 1) / npieces);
            const uint64_t b0 = in->blob_off[plan.chunk_c0[k0]], b1 = in->blob_off[plan.chunk_c0[k1]];
            COPY((uint8_t*)e->d_bytes.p + b0, in->bytes + b0, b1 - b0, cudaMemcpyHostToDevice, e->s_in);
            cudaEvent_t ev = pool_event(e, evi++);
            CK(cudaEventRecord(ev, e->s_in));
            for (uint32_t k = k0; k < k1; k++) ev_in[k] = ev;
        }
    }
    TierTable tt;
    memset(&tt, 0, sizeof tt);
    PlanOut po;
    memset(&po, 0, sizeof po);
    po.status = e->dh.status;
    Dec2Args a;
    memset(&a, 0, sizeof a);
    a.blob_off = e->dh.blob_off; a.bytes = (uint8_t*)e->d_bytes.p;
    a.res_off = e->dh.res_off; a.atom_off = e->dh.atom_off; a.title_off = e->dh.title_off;
    a.seg_off = e->dh.seg_off;
    a.res_type = (uint8_t*)e->d_res_type.p; a.bfactor = (float*)e->d_bfactor.p; a.xyz = (float*)e->d_xyz.p;
    a.titles = out->titles ? (char*)e->d_titles.p : nullptr; a.meta = (fcz_chain_meta*)e->d_meta.p;
    a.status = e->dh.status; a.tables = e->d_tables; a.use_alt = e->opts.use_alt_atom_order;
    static const bool trace = getenv("FCZ_TRACE_HOST") != nullptr;
    std::vector<cudaEvent_t> tr_ev;  // debug timeline: start, then per chunk (kernels done, D2H done)
    if (trace) {
        tr_ev.resize(2 * nchunks + 2);
        for (auto& ev : tr_ev) cudaEventCreate(&ev);
        cudaEventRecord(tr_ev[0], e->stream);
        cudaEventRecord(tr_ev[1], e->s_in);
    }
    // kernels of every chunk first (they only wait for the blobs), then the results back, with at most d2h_depth
    // coordinate copies queued when that is non-zero (see the note at FCZ_CHUNK_BYTES).
    std::vector<cudaEvent_t> ev_k(nchunks), ev_out(nchunks);
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t c0 = subs[k].c0, c1 = subs[k].c1;
        CK(cudaStreamWaitEvent(e->stream, ev_in[k], 0));
        if (c1 > c0) {
            k_dec_plan<<<(c1 - c0 + 7) / 8, 256, 0, e->stream>>>(c0, c1, e->dh.blob_off, (uint8_t*)e->d_bytes.p, e->d_tables, tt, po, 1);
            e->launches++;
            ProfSpan ps(e, FCZ_PROF_DECODE);
            if ((rc = dec2_launch(e, a, subs[k], e->dh.list))) return rc;
        }
        ev_k[k] = pool_event(e, evi++);
        ev_out[k] = pool_event(e, evi++);
        CK(cudaEventRecord(ev_k[k], e->stream));
        if (trace) cudaEventRecord(tr_ev[2 + 2 * k], e->stream);
    }
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t c0 = subs[k].c0, c1 = subs[k].c1;
        if (e->d2h_depth && k >= e->d2h_depth) CK(cudaEventSynchronize(ev_out[k - e->d2h_depth]));
        CK(cudaStreamWaitEvent(e->s_out, ev_k[k], 0));
        const uint64_t a0 = out->atom_off[c0], a1 = out->atom_off[c1];
        COPY(out->xyz + 3ull * a0, (float*)e->d_xyz.p + 3ull * a0, 12ull * (a1 - a0), cudaMemcpyDeviceToHost, e->s_out);
        CK(cudaEventRecord(ev_out[k], e->s_out));
        if (k + 1u == nchunks) {  // the small per-residue / per-chain arrays come back whole, after the last chunk's kernels
            COPY(out->res_type, e->d_res_type.p, n_res, cudaMemcpyDeviceToHost, e->s_out);
            COPY(out->bfactor, e->d_bfactor.p, 4ull * n_res, cudaMemcpyDeviceToHost, e->s_out);
            if (out->titles) COPY(out->titles, e->d_titles.p, n_title, cudaMemcpyDeviceToHost, e->s_out);
            COPY(out->meta, e->d_meta.p, sizeof(fcz_chain_meta) * (uint64_t)n, cudaMemcpyDeviceToHost, e->s_out);
        }
        if (trace) cudaEventRecord(tr_ev[3 + 2 * k], e->s_out);
    }
    if (out->status) {
        cudaEvent_t ev_done = pool_event(e, evi++);
        CK(cudaEventRecord(ev_done, e->stream));
        CK(cudaStreamWaitEvent(e->s_out, ev_done, 0));
        COPY(out->status, e->dh.status, 4ull * n, cudaMemcpyDeviceToHost, e->s_out);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->s_out));
    CK(cudaStreamSynchronize(e->stream));
    if (trace) {
        float t_in = 0;
        cudaEventElapsedTime(&t_in, tr_ev[0], tr_ev[1]);
        fprintf(stderr, "[decode_host trace] %u chunks; inputs landed at %.3f ms\n", nchunks, t_in);
        for (uint32_t k = 0; k < nchunks; k++) {
            float tk = 0, td = 0;
            cudaEventElapsedTime(&tk, tr_ev[0], tr_ev[2 + 2 * k]);
            cudaEventElapsedTime(&td, tr_ev[0], tr_ev[3 + 2 * k]);
            fprintf(stderr, "[decode_host trace] chunk %2u kernels done %.3f  d2h done %.3f\n", k, tk, td);
        }
        for (auto ev : tr_ev) cudaEventDestroy(ev);
    }
    return FCZ_OK;
}

static int decode_plan_device(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals) {
    const uint32_t n = in->n_chains;
    int rc;
    if ((rc = plan_buffers(e, n))) return rc;
    if ((rc = ensure(e, e->d_seg_off, 4ull * (n + 1)))) return rc;
    TierTable tt;
    memset(&tt, 0, sizeof tt);
    PlanOut po;
    po.v0 = (uint32_t*)e->v0.p; po.v1 = (uint32_t*)e->v1.p; po.v2 = (uint32_t*)e->v2.p; po.v3 = (uint32_t*)e->v3.p;
    po.status = out->status ? out->status : (int32_t*)e->status.p;
    po.tier_count = e->d_counters;
    po.tier_list = (uint32_t*)e->tier_list.p;
    po.enc_desc = nullptr;
    if (n) {
        k_dec_plan<<<(n + 7) / 8, 256, 0, e->stream>>>(0u, n, in->blob_off, in->bytes, e->d_tables, tt, po, 0);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 4;
    sa.in[0] = po.v0; sa.out[0] = out->res_off; sa.out64[0] = 0;
    sa.in[1] = po.v1; sa.out[1] = out->atom_off; sa.out64[1] = 1;
    sa.in[2] = po.v2; sa.out[2] = out->title_off; sa.out64[2] = 0;
    sa.in[3] = po.v3; sa.out[3] = e->d_seg_off.p; sa.out64[3] = 0;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = ensure(e, e->d_dec_list, 4ull * FCZ_DEC_TIERS * n + 16))) return rc;
    static const bool use_desc = [] { const char* v = getenv("FCZ_DEC_DESC"); return !v || atoi(v) != 0; }();  // (A/B knob)
    if (use_desc && (rc = ensure(e, e->d_dec_desc, sizeof(DecDesc) * (size_t)FCZ_DEC_TIERS * n + 64))) return rc;
    e->dec_desc_valid = use_desc;
    k_plan_chunks<<<n / 256 + 1, 256, 0, e->stream>>>(n, e->dec_sub_res, out->res_off, (uint32_t*)e->d_seg_off.p, out->atom_off, in->blob_off,
                                                  po.status, (uint32_t*)e->d_dec_list.p, (uint32_t*)e->d_submax.p, e->h_bounds, out->title_off,
                                                  use_desc ? (DecDesc*)e->d_dec_desc.p : nullptr);
    e->launches++;
    if ((rc = fetch_plan(e))) return rc;
    totals->n_res = e->h_totals[0];
    totals->n_atoms = e->h_totals[1];
    totals->n_title_bytes = e->h_totals[2];
    totals->n_blob_bytes = 0;
    return FCZ_OK;
}

static int decode_device(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out) {
    // uses the sub-batch bounds left in pinned memory by the last decode_plan_device on this engine
    if (e->h_totals[0] > out->res_cap || e->h_totals[1] > out->atom_cap || (out->titles && e->h_totals[2] > out->title_cap))
        return fail(e, FCZ_E_CAPACITY, "decode output capacity too small (need %llu residues, %llu atoms, %llu title bytes)",
                    (unsigned long long)e->h_totals[0], (unsigned long long)e->h_totals[1], (unsigned long long)e->h_totals[2]);
    const uint32_t nsub = e->h_bounds[0], n = in->n_chains;
    std::vector<Dec2Sub> subs(nsub);
    uint32_t per = nsub ? e->h_bounds[4] - e->h_bounds[1] : 0u;  // chains per sub-batch (the last may be shorter)
    for (uint32_t k = 0; k < nsub; k++) {
        const uint32_t* b = e->h_bounds + 1 + 3 * k;
        const uint32_t* m = e->h_bounds + 1 + 3 * (nsub + 1) + 5 * FCZ_DEC_TIERS * k;
        Dec2Sub& sb = subs[k];
        memset(&sb, 0, sizeof sb);
        sb.c0 = b[0]; sb.c1 = b[3]; sb.r0 = b[1]; sb.r1 = b[4]; sb.s0 = b[2]; sb.s1 = b[5];
        for (int t = 0; t < FCZ_DEC_TIERS; t++) {
            Dec2Tier& tr = sb.tier[t];
            tr.count = m[5 * t]; tr.max_L = m[5 * t + 1]; tr.max_anchor = m[5 * t + 2]; tr.max_blob = m[5 * t + 3]; tr.max_atoms = m[5 * t + 4];
            tr.list_off = (size_t)t * n + (size_t)k * per;
        }
    }
    int rc;
    if ((rc = dec2_workspace(e, subs.data(), subs.size()))) return rc;
    Dec2Args a;
    memset(&a, 0, sizeof a);
    a.blob_off = in->blob_off; a.bytes = in->bytes;
    a.res_off = out->res_off; a.atom_off = out->atom_off; a.title_off = out->title_off; a.seg_off = (uint32_t*)e->d_seg_off.p;
    a.res_type = out->res_type; a.bfactor = out->bfactor; a.xyz = out->xyz; a.titles = out->titles; a.meta = out->meta;
    a.status = out->status ? out->status : (int32_t*)e->status.p;
    a.tables = e->d_tables; a.use_alt = e->opts.use_alt_atom_order;
    {
        ProfSpan ps(e, FCZ_PROF_DECODE);
        for (uint32_t k = 0; k < nsub; k++)
            if ((rc = dec2_launch(e, a, subs[k], (const uint32_t*)e->d_dec_list.p, e->dec_desc_valid ? (const DecDesc*)e->d_dec_desc.p : nullptr))) return rc;
    }
    CK(cudaGetLastError());
    return FCZ_OK;
}

extern "C" int fcz_decode_plan(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out, fcz_sizes* totals) {
    if (!e || !in || !out || !totals) return FCZ_E_ARG;
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    int rc;
    if (in->mem == FCZ_MEM_DEVICE) {
        rc = decode_plan_device(e, in, out, totals);
        return rc;
    }
    (void)rc;
    return decode_plan_host(e, in, out, totals);
}

extern "C" int fcz_decode_batch(fcz_engine* e, const fcz_blob_batch* in, fcz_chain_batch* out) {
    if (!e || !in || !out) return FCZ_E_ARG;
    if (in->mem != out->mem) return fail(e, FCZ_E_ARG, "input and output batches must live in the same memory space");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    int rc;
    if (in->mem == FCZ_MEM_DEVICE) return decode_device(e, in, out);
    (void)n; (void)rc;
    return decode_host(e, in, out);
}

// --------------------------------------------------------------------- fused decode -> PDB text (host blobs in, host text out)
// What `foldcomp decompress` does per entry (src/main.cpp:612-689: read + decompress + writeAtomCoordinatesToPDB), for a
// batch: blobs go up once, the decoded coordinates never leave the GPU, only the text comes back.
extern "C" int fcz_decode_to_pdb_plan(fcz_engine* e, const fcz_blob_batch* in, fcz_text_batch* out, uint64_t* total_bytes) {
    if (!e || !in || !out || !total_bytes) return FCZ_E_ARG;
    if (in->mem != FCZ_MEM_HOST || out->mem != FCZ_MEM_HOST) return fail(e, FCZ_E_ARG, "fcz_decode_to_pdb_* take host-memory batches");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    e->pdb.valid = false;
    int rc;
    const uint64_t nb = in->blob_off[n];
    H2D(e->d_blob_off, in->blob_off, 8ull * (n + 1));
    H2D(e->d_bytes, in->bytes, nb);
    if ((rc = ensure(e, e->d_res_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_atom_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_title_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_status, 4ull * n + 4))) return rc;
    if ((rc = ensure(e, e->d_text_off, 8ull * (n + 1)))) return rc;
    fcz_blob_batch bin;
    memset(&bin, 0, sizeof bin);
    bin.n_chains = n; bin.mem = FCZ_MEM_DEVICE; bin.blob_off = (uint64_t*)e->d_blob_off.p; bin.bytes = (uint8_t*)e->d_bytes.p;
    fcz_chain_batch cb;
    memset(&cb, 0, sizeof cb);
    cb.n_chains = n; cb.mem = FCZ_MEM_DEVICE;
    cb.res_off = (uint32_t*)e->d_res_off.p; cb.atom_off = (uint64_t*)e->d_atom_off.p; cb.title_off = (uint32_t*)e->d_title_off.p;
    cb.status = (int32_t*)e->d_status.p;
    fcz_sizes tot;
    if ((rc = decode_plan_device(e, &bin, &cb, &tot))) return rc;
    if ((rc = ensure(e, e->d_res_type, tot.n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_bfactor, 4ull * tot.n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_xyz, 12ull * tot.n_atoms + 16))) return rc;
    if ((rc = ensure(e, e->d_titles, tot.n_title_bytes + 16))) return rc;
    if ((rc = ensure(e, e->d_meta, sizeof(fcz_chain_meta) * (uint64_t)n + 16))) return rc;
    cb.res_type = (uint8_t*)e->d_res_type.p; cb.bfactor = (float*)e->d_bfactor.p; cb.xyz = (float*)e->d_xyz.p;
    cb.titles = (char*)e->d_titles.p; cb.meta = (fcz_chain_meta*)e->d_meta.p;
    cb.res_cap = tot.n_res; cb.atom_cap = tot.n_atoms; cb.title_cap = tot.n_title_bytes;
    if ((rc = decode_device(e, &bin, &cb))) return rc;
    if ((rc = pdb_plan_dev(e, staged_chains(e), n, tot.n_res, (uint64_t*)e->d_text_off.p, out->text_off))) return rc;
    if (out->status) {
        CK(cudaMemcpyAsync(out->status, e->d_status.p, 4ull * n, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    *total_bytes = e->pdb.total_bytes;
    return FCZ_OK;
}

extern "C" int fcz_decode_to_pdb_batch(fcz_engine* e, const fcz_blob_batch* in, fcz_text_batch* out) {
    if (!e || !in || !out) return FCZ_E_ARG;
    if (in->mem != FCZ_MEM_HOST || out->mem != FCZ_MEM_HOST) return fail(e, FCZ_E_ARG, "fcz_decode_to_pdb_* take host-memory batches");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    if (!e->pdb.valid || e->pdb.n != n) return fail(e, FCZ_E_ARG, "fcz_decode_to_pdb_batch needs a preceding fcz_decode_to_pdb_plan on the same batch");
    if (e->pdb.total_bytes > out->bytes_cap)
        return fail(e, FCZ_E_CAPACITY, "text needs %llu bytes, capacity %llu", (unsigned long long)e->pdb.total_bytes, (unsigned long long)out->bytes_cap);
    e->pdb.valid = false;
    return pdb_emit_to_host(e, staged_chains(e), n, out->text_off, out->bytes);
}

// ---------------------------------------------------------------------------------- continuised angles (get_data)
extern "C" int fcz_unpack_angles_batch(fcz_engine* e, const fcz_blob_batch* in, uint64_t* res_off, float* angles, uint64_t res_cap,
                                       uint64_t* total_res) {
    if (!e || !in || !res_off || !total_res) return FCZ_E_ARG;
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    int rc;
    if ((rc = plan_buffers(e, n))) return rc;
    const bool host = in->mem == FCZ_MEM_HOST;
    ExtractArgs x;
    memset(&x, 0, sizeof x);
    if (host) {
        H2D(e->d_blob_off, in->blob_off, 8ull * (n + 1));
        H2D(e->d_bytes, in->bytes, in->blob_off[n]);
        if ((rc = ensure(e, e->d_text_off, 8ull * (n + 1)))) return rc;
        x.blob_off = (uint64_t*)e->d_blob_off.p; x.bytes = (uint8_t*)e->d_bytes.p;
    } else {
        x.blob_off = in->blob_off; x.bytes = in->bytes;
    }
    x.tt = e->d_text_tables; x.type = 1; x.digits = 1; x.n = n;  // type 1 sizes = residues per blob
    x.text_bytes = (uint32_t*)e->v0.p; x.status = (int32_t*)e->status.p;
    uint64_t* d_res_off = host ? (uint64_t*)e->d_text_off.p : res_off;
    if (n) {
        k_extract_plan<<<(n + 255) / 256, 256, 0, e->stream>>>(x);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 1;
    sa.in[0] = x.text_bytes; sa.out[0] = d_res_off; sa.out64[0] = 1;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = fetch_plan(e))) return rc;
    const uint64_t total = e->h_totals[0];
    *total_res = total;
    if (!angles || total > res_cap) return total > res_cap && angles ? fail(e, FCZ_E_CAPACITY, "angles need %llu residues of capacity", (unsigned long long)total) : FCZ_OK;
    if (host && (rc = ensure(e, e->d_text, 24ull * total + 64))) return rc;
    AnglesArgs a;
    a.blob_off = x.blob_off; a.bytes = x.bytes; a.status = x.status; a.res_off = d_res_off; a.n = n;
    a.angles = host ? (float*)e->d_text.p : angles;
    if (n) {
        k_unpack_angles<<<n, 128, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    if (host) {
        CK(cudaMemcpyAsync(res_off, d_res_off, 8ull * (n + 1), cudaMemcpyDeviceToHost, e->stream));
        if (total) CK(cudaMemcpyAsync(angles, e->d_text.p, 24ull * total, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return FCZ_OK;
}

// ---------------------------------------------------------------------------------- backbone angles before quantisation
extern "C" int fcz_backbone_angles_batch(fcz_engine* e, const fcz_chain_batch* in, float* angles) {
    if (!e || !in || !angles) return FCZ_E_ARG;
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    int rc;
    DevChains d;
    const bool host = in->mem == FCZ_MEM_HOST;
    uint64_t n_res = in->res_cap;
    if (host) {
        n_res = in->res_off[n];
        H2D(e->d_res_off, in->res_off, 4ull * (n + 1));
        H2D(e->d_atom_off, in->atom_off, 8ull * (n + 1));
        H2D(e->d_res_type, in->res_type, n_res);
        H2D(e->d_xyz, in->xyz, 12ull * in->atom_off[n]);
        d = staged_chains(e);
        if ((rc = ensure(e, e->d_text, 24ull * n_res + 64))) return rc;
    } else if ((rc = upload_chains(e, in, &d))) return rc;
    if ((rc = ensure(e, e->ws_aoff, 4ull * (n_res + n + 1)))) return rc;
    RawAnglesArgs a;
    a.res_off = d.res_off; a.atom_off = d.atom_off; a.res_type = d.res_type; a.xyz = d.xyz; a.tables = e->d_tables;
    a.aoff = (uint32_t*)e->ws_aoff.p; a.angles = host ? (float*)e->d_text.p : angles; a.n = n;
    if (n) {
        k_raw_angles<<<n, 128, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    if (host) {
        if (n_res) CK(cudaMemcpyAsync(angles, e->d_text.p, 24ull * n_res, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return FCZ_OK;
}

// ---------------------------------------------------------------------------------- check (Foldcomp::checkValidity)
extern "C" int fcz_check_batch(fcz_engine* e, const fcz_blob_batch* in, int32_t* read_status, int32_t* validity) {
    if (!e || !in) return FCZ_E_ARG;
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    int rc;
    const bool host = in->mem == FCZ_MEM_HOST;
    CheckArgs a;
    memset(&a, 0, sizeof a);
    a.n = n;
    if (host) {
        H2D(e->d_blob_off, in->blob_off, 8ull * (n + 1));
        H2D(e->d_bytes, in->bytes, in->blob_off[n]);
        if ((rc = ensure(e, e->d_status, 8ull * n + 8))) return rc;
        a.blob_off = (uint64_t*)e->d_blob_off.p; a.bytes = (uint8_t*)e->d_bytes.p;
        a.read_status = (int32_t*)e->d_status.p; a.validity = (int32_t*)e->d_status.p + n;
    } else {
        a.blob_off = in->blob_off; a.bytes = in->bytes; a.read_status = read_status; a.validity = validity;
    }
    if (n) {
        k_check<<<(n + 7) / 8, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    if (host) {
        if (read_status && n) CK(cudaMemcpyAsync(read_status, a.read_status, 4ull * n, cudaMemcpyDeviceToHost, e->stream));
        if (validity && n) CK(cudaMemcpyAsync(validity, a.validity, 4ull * n, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
    }
    return FCZ_OK;
}

// ---------------------------------------------------------------------------------- PDB text in (section 8 f3)
static void parse_args(fcz_engine* e, const uint64_t* text_off, const char* text, uint32_t n, ParseArgs* a) {
    memset(a, 0, sizeof *a);
    a->text_off = text_off; a->text = text; a->n = n; a->pt = e->d_parse_tables;
    a->n_lines = (uint32_t*)e->v0.p; a->line_off = (uint64_t*)e->p_line_off.p;
    a->lines = (uint32_t*)e->p_lines.p; a->rstart = (uint32_t*)e->p_rstart.p; a->raw = (RawAtom*)e->p_raw.p;
    a->scratch = (uint32_t*)e->p_scratch.p;
    a->v_res = (uint32_t*)e->v1.p; a->v_atoms = (uint32_t*)e->v2.p;
}

// lines -> workspace -> per-entry plan -> offsets; two stream syncs (the workspace and the outputs are sized by the text)
static int parse_plan_dev(fcz_engine* e, const uint64_t* d_text_off, const char* d_text, uint32_t n, uint32_t* d_res_off,
                          uint64_t* d_atom_off, int32_t* d_status, fcz_sizes* totals) {
    int rc;
    e->parse.valid = false;
    if (!e->d_parse_tables) {
        ParseTables h;
        build_parse_tables(&h);
        CK(cudaMalloc(&e->d_parse_tables, sizeof(ParseTables)));
        CK(cudaMemcpyAsync(e->d_parse_tables, &h, sizeof(ParseTables), cudaMemcpyHostToDevice, e->stream));
        CK(cudaStreamSynchronize(e->stream));  // h is a stack object
    }
    if ((rc = plan_buffers(e, n))) return rc;
    if ((rc = ensure(e, e->p_line_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->p_scratch, 32ull * n + 32))) return rc;
    ParseArgs a;
    parse_args(e, d_text_off, d_text, n, &a);
    if (n) {
        k_parse_lines<<<n, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    ScanArgs sa;
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 1;
    sa.in[0] = a.n_lines; sa.out[0] = e->p_line_off.p; sa.out64[0] = 1;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = fetch_plan(e))) return rc;
    const uint64_t total_lines = e->h_totals[0];
    if ((rc = ensure(e, e->p_lines, 4ull * (total_lines + n + 1)))) return rc;
    if ((rc = ensure(e, e->p_rstart, 4ull * (total_lines + n + 1)))) return rc;
    if ((rc = ensure(e, e->p_raw, 2ull * sizeof(RawAtom) * total_lines + 64))) return rc;
    parse_args(e, d_text_off, d_text, n, &a);
    a.status = d_status;
    if (n) {
        k_parse_plan<<<n, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    memset(&sa, 0, sizeof sa);
    sa.n = n; sa.narr = 2;
    sa.in[0] = a.v_res; sa.out[0] = d_res_off; sa.out64[0] = 0;
    sa.in[1] = a.v_atoms; sa.out[1] = d_atom_off; sa.out64[1] = 1;
    if ((rc = run_scan(e, sa))) return rc;
    if ((rc = fetch_plan(e))) return rc;
    if (e->h_totals[0] > 0xFFFFFFFFull) return fail(e, FCZ_E_LIMIT, "a parsed batch holds %llu residues (the canonical layout indexes residues with 32 bits)", (unsigned long long)e->h_totals[0]);
    if (totals) { totals->n_res = e->h_totals[0]; totals->n_atoms = e->h_totals[1]; totals->n_title_bytes = 0; totals->n_blob_bytes = 0; }
    e->parse.n = n; e->parse.n_res = e->h_totals[0]; e->parse.n_atoms = e->h_totals[1]; e->parse.status = d_status; e->parse.valid = true;
    return FCZ_OK;
}

static int parse_emit_dev(fcz_engine* e, const uint64_t* d_text_off, const char* d_text, uint32_t n, const uint32_t* d_res_off,
                          const uint64_t* d_atom_off, uint8_t* res_type, float* bfactor, float* xyz, fcz_chain_meta* meta) {
    ParseArgs a;
    parse_args(e, d_text_off, d_text, n, &a);
    a.status = e->parse.status;
    a.res_off = d_res_off; a.atom_off = d_atom_off; a.res_type = res_type; a.bfactor = bfactor; a.xyz = xyz; a.meta = meta;
    if (n) {
        k_parse_emit<<<n, 256, 0, e->stream>>>(a);
        e->launches++;
    }
    CK(cudaGetLastError());
    return FCZ_OK;
}

extern "C" int fcz_parse_pdb_plan(fcz_engine* e, const fcz_text_batch* in, fcz_chain_batch* out, fcz_sizes* totals) {
    if (!e || !in || !out || !totals) return FCZ_E_ARG;
    if (in->mem != FCZ_MEM_DEVICE || out->mem != FCZ_MEM_DEVICE) return fail(e, FCZ_E_ARG, "fcz_parse_pdb_plan works on device memory (fcz_encode_pdb_text_batch takes host text)");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    int rc;
    int32_t* st = out->status;
    if (!st) { if ((rc = ensure(e, e->d_status, 4ull * n + 4))) return rc; st = (int32_t*)e->d_status.p; }
    return parse_plan_dev(e, in->text_off, in->bytes, n, out->res_off, out->atom_off, st, totals);
}

extern "C" int fcz_parse_pdb_batch(fcz_engine* e, const fcz_text_batch* in, fcz_chain_batch* out) {
    if (!e || !in || !out) return FCZ_E_ARG;
    if (in->mem != FCZ_MEM_DEVICE || out->mem != FCZ_MEM_DEVICE) return fail(e, FCZ_E_ARG, "fcz_parse_pdb_batch works on device memory");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    if (!e->parse.valid || e->parse.n != n) return fail(e, FCZ_E_ARG, "fcz_parse_pdb_batch needs a preceding fcz_parse_pdb_plan on the same batch");
    if (e->parse.n_res > out->res_cap || e->parse.n_atoms > out->atom_cap)
        return fail(e, FCZ_E_CAPACITY, "parsed batch needs %llu residues / %llu atoms of capacity", (unsigned long long)e->parse.n_res, (unsigned long long)e->parse.n_atoms);
    e->parse.valid = false;
    return parse_emit_dev(e, in->text_off, in->bytes, n, out->res_off, out->atom_off, out->res_type, out->bfactor, out->xyz, out->meta);
}

extern "C" int fcz_encode_pdb_text_batch(fcz_engine* e, const fcz_text_batch* in, const uint32_t* title_off, const char* titles,
                                         fcz_blob_batch* out, uint64_t* total_bytes) {
    if (!e || !in || !out || !total_bytes || !title_off) return FCZ_E_ARG;
    if (in->mem != FCZ_MEM_HOST || out->mem != FCZ_MEM_HOST) return fail(e, FCZ_E_ARG, "fcz_encode_pdb_text_batch takes host text and returns host blobs");
    CK(cudaSetDevice(e->device));
    const uint32_t n = in->n_chains;
    out->n_chains = n;
    int rc;
    const uint64_t n_text = in->text_off[n], n_title = title_off[n];
    // text up: the offsets first (the line count needs them), the bytes in one copy
    H2D(e->d_text_off, in->text_off, 8ull * (n + 1));
    H2D(e->d_text, in->bytes, n_text);
    H2D(e->d_title_off, title_off, 4ull * (n + 1));
    H2D(e->d_titles, titles, n_title);
    if ((rc = ensure(e, e->d_res_off, 4ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_atom_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_status, 8ull * n + 8))) return rc;
    if ((rc = ensure(e, e->d_meta, sizeof(fcz_chain_meta) * (uint64_t)n + 32))) return rc;
    int32_t* d_pstat = (int32_t*)e->d_status.p;       // parser status
    int32_t* d_estat = (int32_t*)e->d_status.p + n;   // encoder status
    fcz_sizes tot;
    if ((rc = parse_plan_dev(e, (const uint64_t*)e->d_text_off.p, (const char*)e->d_text.p, n, (uint32_t*)e->d_res_off.p,
                             (uint64_t*)e->d_atom_off.p, d_pstat, &tot))) return rc;
    if ((rc = ensure(e, e->d_res_type, tot.n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_bfactor, 4ull * tot.n_res + 16))) return rc;
    if ((rc = ensure(e, e->d_xyz, 12ull * tot.n_atoms + 16))) return rc;
    if (n) CK(cudaMemsetAsync(e->d_meta.p, 0, sizeof(fcz_chain_meta) * (uint64_t)n, e->stream));  // entries that failed to parse
    e->parse.valid = false;
    if ((rc = parse_emit_dev(e, (const uint64_t*)e->d_text_off.p, (const char*)e->d_text.p, n, (uint32_t*)e->d_res_off.p,
                             (uint64_t*)e->d_atom_off.p, (uint8_t*)e->d_res_type.p, (float*)e->d_bfactor.p, (float*)e->d_xyz.p,
                             (fcz_chain_meta*)e->d_meta.p))) return rc;
    // encode the device-resident chains
    const uint64_t bound = fcz_encode_bound(n, tot.n_res, tot.n_atoms, n_title, e->opts.anchor_threshold);
    if ((rc = ensure(e, e->d_blob_off, 8ull * (n + 1)))) return rc;
    if ((rc = ensure(e, e->d_bytes, bound + 64))) return rc;
    fcz_chain_batch cb;
    memset(&cb, 0, sizeof cb);
    cb.n_chains = n; cb.mem = FCZ_MEM_DEVICE;
    cb.res_off = (uint32_t*)e->d_res_off.p; cb.atom_off = (uint64_t*)e->d_atom_off.p; cb.title_off = (uint32_t*)e->d_title_off.p;
    cb.res_type = (uint8_t*)e->d_res_type.p; cb.bfactor = (float*)e->d_bfactor.p; cb.xyz = (float*)e->d_xyz.p;
    cb.titles = (char*)e->d_titles.p; cb.meta = (fcz_chain_meta*)e->d_meta.p;
    fcz_blob_batch bb;
    memset(&bb, 0, sizeof bb);
    bb.n_chains = n; bb.mem = FCZ_MEM_DEVICE;
    bb.blob_off = (uint64_t*)e->d_blob_off.p; bb.bytes = (uint8_t*)e->d_bytes.p; bb.status = d_estat; bb.bytes_cap = bound;
    uint64_t total = 0;
    if ((rc = encode_device(e, &cb, &bb, &total))) return rc;
    *total_bytes = total;
    if (total > out->bytes_cap) return fail(e, FCZ_E_CAPACITY, "encode needs %llu bytes, capacity %llu", (unsigned long long)total, (unsigned long long)out->bytes_cap);
    CK(cudaMemcpyAsync(out->blob_off, e->d_blob_off.p, 8ull * (n + 1), cudaMemcpyDeviceToHost, e->stream));
    if (total) CK(cudaMemcpyAsync(out->bytes, e->d_bytes.p, total, cudaMemcpyDeviceToHost, e->stream));
    e->h_parse_status.resize(2ull * n + 2);
    if (n) CK(cudaMemcpyAsync(e->h_parse_status.data(), e->d_status.p, 8ull * n, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (out->status)
        for (uint32_t c = 0; c < n; c++) out->status[c] = e->h_parse_status[c] ? e->h_parse_status[c] : e->h_parse_status[n + c];
    return FCZ_OK;
}
