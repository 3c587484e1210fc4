// semloss_tc.cuh -- the two dense contractions of the training-side semantic loss (reference train.py:150 and its
// autograd backward) on the 5th-generation tensor cores (tcgen05.mma, accumulators in tensor memory), fp32-accurate
// through the x = hi + lo TF32 split (three products, small terms first; ~2^-21 of an fp32 FMA chain).
//
//   k_sim_tc     sim = (gt / |gt|) @ lut1^T  for 128 pixels x all K codebook rows per tile, accumulated in TMEM over
//                chunks of KC codebook-width elements, and -- while the row is still in tensor memory -- everything
//                train.py:152-160 does with it: row max / arg-max / label bits, softmax(t sim) entropy, the loss
//                partial sums and d loss / d sim.  The similarity matrix itself never reaches HBM; 1/|gt| is
//                accumulated from the operand tiles as they are staged (no separate pass over gt).
//   k_dlut_tc    dlut1^T[d, k] = sum_p gt[p, d] * dsim[p, k]: a 128-wide slice of the codebook width per CTA, all K
//                codebook rows as accumulator columns, the pixel axis as the reduction, split over the CTAs; partial
//                sums leave TMEM through red.global.add.f32.
//
// Both kernels: one persistent CTA per SM (256 threads) that owns the SM's tensor memory (512 columns); operands in
// shared memory in the K-major SWIZZLED canonical layout -- a row is one chunk of KC reduction elements (128 B with
// SWIZZLE_128B at KC = 32, 64 B with SWIZZLE_64B at KC = 16), rows are contiguous, and the 16-byte piece c of row r
// sits at piece c ^ x(r), x(r) = r % 8 | (r / 2) % 4 (recipe verified in profiles/micro/tc05_probe_sw.cu).  The XOR is
// what lets a warp read 128 contiguous bytes of a source row AND store them without bank conflicts: with the
// unswizzled layout the loads had to be spread over 8 rows per instruction, which saturated the L1 data pipe --
// two stages, each a TF32 hi image and a lo image; a stage is refilled as soon as the MMAs that read it have
// committed (mbarrier), so the tensor pipe always has the next chunk queued behind the running one.  Operands that
// come from fp32 tensors pass through registers (split, 4x4 transposes where the source is contiguous along the
// wrong axis); the codebook operand of k_sim_tc is pre-split once per call (k_build_wimg) and fetched with one bulk
// async copy (cp.async.bulk + mbarrier complete_tx) per image.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc5 {

constexpr int THREADS = 256;
constexpr int MAX_STAGES = 4;                               // operand stages (each: hi + lo images of both operands)
__host__ __device__ constexpr int row_bytes(int kc) { return 4 * kc; }                      // one operand row of one chunk
__host__ __device__ constexpr int swz(int kc, int r) { return kc == 32 ? (r & 7) : ((r >> 1) & 3); }
// byte offset of the 16-byte piece c (reduction elements 4c .. 4c+3) of row r inside an operand image
__host__ __device__ constexpr int piece_off(int kc, int r, int c) { return r * row_bytes(kc) + ((c ^ swz(kc, r)) << 4); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int KC>
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address (image bases are 1024-byte aligned)
    d |= (uint64_t)1 << 16;                                 // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(((8 * row_bytes(KC)) >> 4) & 0x3fff) << 32;   // stride byte offset: 8-row groups are contiguous
    d |= (uint64_t)1 << 46;                                 // descriptor version 1 (sm_100)
    d |= (uint64_t)(KC == 32 ? 2 : 4) << 61;                // SWIZZLE_128B | SWIZZLE_64B
    return d;
}
__device__ __forceinline__ uint32_t instr_desc(int m, int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);   // F32 += TF32 x TF32, K-major
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate));
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("{\n.reg .b64 t;\nmbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1;\n}\n" ::"r"(bar), "r"(bytes) : "memory");
}
// one bulk asynchronous copy global -> shared (multiple of 16 bytes, 16-byte aligned both sides); completion is
// signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// columns [c, c + cnt) of this warp's 32 TMEM lanes, cnt = 32 or 16 (warp-uniform)
__device__ __forceinline__ void ld_cols(uint32_t taddr, int cnt, uint32_t (&v)[32])
{
    if (cnt >= 32) ld32(taddr, v); else ld16(taddr, v);
    ld_wait();
}
// without the wait: the caller overlaps the load with arithmetic on the previous block
__device__ __forceinline__ void ld_cols_nowait(uint32_t taddr, int cnt, uint32_t (&v)[32])
{
    if (cnt >= 32) ld32(taddr, v); else ld16(taddr, v);
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Calls f(v, c0, cnt, full) for the 32-column blocks of [cbeg, cend) (the last one may be 16 wide), with the TMEM load
// of block b + 1 in flight while block b is processed.  full = every column of the block is a real codebook row.
template <typename F>
__device__ __forceinline__ void for_blocks(uint32_t lane_base, int cbeg, int cend, int K, F&& f)
{
    uint32_t va[32], vb[32];
    if (cbeg >= cend) return;
    ld_cols_nowait(lane_base + (uint32_t)cbeg, min(32, cend - cbeg), va);
    for (int c0 = cbeg; c0 < cend; c0 += 64) {
        const int c1 = c0 + 32, c2 = c0 + 64;
        ld_wait();
        if (c1 < cend) ld_cols_nowait(lane_base + (uint32_t)c1, min(32, cend - c1), vb);
        f(va, c0, min(32, cend - c0), c0 + 32 <= K);
        if (c1 < cend) {
            ld_wait();
            if (c2 < cend) ld_cols_nowait(lane_base + (uint32_t)c2, min(32, cend - c2), va);
            f(vb, c1, min(32, cend - c1), c1 + 32 <= K);
        }
    }
}
// the same without the second register buffer (kernels with 16 warps hide the TMEM latency with other warps)
template <typename F>
__device__ __forceinline__ void for_blocks1(uint32_t lane_base, int cbeg, int cend, int K, F&& f)
{
    uint32_t v[32];
    for (int c0 = cbeg; c0 < cend; c0 += 32) {
        const int cnt = min(32, cend - c0);
        ld_cols(lane_base + (uint32_t)c0, cnt, v);
        f(v, c0, cnt, c0 + 32 <= K);
    }
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts128f(uint32_t a, float x, float y, float z, float w)
{
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// TF32 hi / lo images of four consecutive reduction elements of one operand row
__device__ __forceinline__ void store_split(uint32_t hi_base, uint32_t lo_base, uint32_t off, float a, float b, float c, float d, bool want_lo)
{
    const uint32_t ha = __float_as_uint(a) & 0xffffe000u, hb = __float_as_uint(b) & 0xffffe000u,
                   hc = __float_as_uint(c) & 0xffffe000u, hd = __float_as_uint(d) & 0xffffe000u;
    sts128u(hi_base + off, ha, hb, hc, hd);
    if (want_lo)
        sts128f(lo_base + off, a - __uint_as_float(ha), b - __uint_as_float(hb), c - __uint_as_float(hc), d - __uint_as_float(hd));
}
// four consecutive floats with the first `nvalid` inside the tensor (zeros behind); `vec` = 16-byte loads are legal
__device__ __forceinline__ float4 ld4(const float* p, int nvalid, bool vec)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nvalid >= 4 && vec) {
        v = __ldg(reinterpret_cast<const float4*>(p));
    } else if (nvalid > 0) {
        v.x = __ldg(p);
        if (nvalid > 1) v.y = __ldg(p + 1);
        if (nvalid > 2) v.z = __ldg(p + 2);
        if (nvalid > 3) v.w = __ldg(p + 3);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------------------------------
// A 128-row operand tile that comes from an fp32 tensor, one chunk of KC reduction elements at a time.
//   TRANS = false: the source is contiguous along the reduction axis, element (r, k) = src[(row0 + r) ld + k0 + k].
//                  A thread owns 16-byte pieces (row, four k); 8 consecutive lanes = 8 consecutive rows (conflict-free
//                  stores), the lane's upper bits walk k: a warp instruction reads 64 contiguous bytes of 8 rows.
//   TRANS = true:  the source is contiguous along the row axis, element (r, k) = src[(k0 + k) ld + row0 + r].
//                  A thread owns a 4 x 4 block (rows 4 lane .. 4 lane + 3, four k = its warp's quad), loaded as four
//                  16-byte vectors (coalesced: a warp reads 512 contiguous bytes per k) and transposed in registers.
// Rows >= row_end and reduction elements >= k_end read as zeros.
// ---------------------------------------------------------------------------------------------------------------------
template <int KC, bool TRANS, int NW = 8>                   // NW = warps of the CTA (8 or 16); the transposed mode uses the first 8
struct RowTile {
    static constexpr int Q = KC / 4;                        // 16-byte pieces per row
    static constexpr int RPU = 32 / Q;                      // rows per warp instruction (direct mode)
    static constexpr int NREG = TRANS ? 4 : 4 * Q / NW;     // float4 registers per thread
    float4 v[NREG];

    // transposed mode: lane = hi * 8 + k4l * 2 + lo -> row quad 8 (warp % 4) + 2 hi + lo, piece 4 (warp / 4) + k4l.
    // A store phase (8 consecutive lanes) then holds 4 pieces x 2 row parities = 8 different swizzled positions, and a
    // load instruction reads 4 source rows x 128 contiguous bytes.
    static __device__ __forceinline__ int t_r4(int tid) { const int lane = tid & 31; return 8 * ((tid >> 5) & 3) + 2 * (lane >> 3) + (lane & 1); }
    static __device__ __forceinline__ int t_k4(int tid) { return 4 * (tid >> 7) + ((tid >> 1) & 3); }

    __device__ __forceinline__ void load(const float* __restrict__ src, int64_t ld, int64_t row0, int64_t row_end,
                                         int64_t k0, int64_t k_end, bool vec, int tid)
    {
        if (TRANS) {
            const int k4 = t_k4(tid);                       // warps >= 4 idle at KC = 16
            const int64_t r = row0 + 4 * t_r4(tid);
            const int nv = (int)min((int64_t)4, row_end - r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t k = k0 + 4 * k4 + j;
                v[j] = (k4 < Q && k < k_end) ? ld4(src + k * ld + r, nv, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const int lane = tid & 31, w = tid >> 5;
#pragma unroll
            for (int it = 0; it < NREG; ++it) {
                const int r = (it * NW + w) * RPU + lane / Q, c = lane % Q;
                const int64_t k = k0 + 4 * c;
                v[it] = (row0 + r < row_end) ? ld4(src + (row0 + r) * ld + k, (int)min((int64_t)4, k_end - k), vec)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    // writes the hi (and lo) images of the loaded chunk; ss[] += squares per owned row (row_of() names them)
    __device__ __forceinline__ void store(uint32_t hi_base, uint32_t lo_base, bool want_lo, int tid, float (&ss)[4])
    {
        if (TRANS) {
            const int k4 = t_k4(tid), r4 = t_r4(tid);
            if (k4 < Q) {
                const float x[4][4] = {{v[0].x, v[1].x, v[2].x, v[3].x}, {v[0].y, v[1].y, v[2].y, v[3].y},
                                       {v[0].z, v[1].z, v[2].z, v[3].z}, {v[0].w, v[1].w, v[2].w, v[3].w}};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    store_split(hi_base, lo_base, (uint32_t)piece_off(KC, 4 * r4 + i, k4), x[i][0], x[i][1], x[i][2], x[i][3], want_lo);
                    ss[i] = fmaf(x[i][0], x[i][0], fmaf(x[i][1], x[i][1], fmaf(x[i][2], x[i][2], fmaf(x[i][3], x[i][3], ss[i]))));
                }
            }
        } else {
            const int lane = tid & 31, w = tid >> 5;
#pragma unroll
            for (int it = 0; it < NREG; ++it) {
                const int r = (it * NW + w) * RPU + lane / Q, c = lane % Q;
                store_split(hi_base, lo_base, (uint32_t)piece_off(KC, r, c), v[it].x, v[it].y, v[it].z, v[it].w, want_lo);
                ss[it] = fmaf(v[it].x, v[it].x, fmaf(v[it].y, v[it].y, fmaf(v[it].z, v[it].z, fmaf(v[it].w, v[it].w, ss[it]))));
            }
        }
    }
    // the row that ss[i] of store() belongs to (-1: none)
    static __device__ __forceinline__ int row_of(int i, int tid)
    {
        if (TRANS) return t_k4(tid) < Q ? 4 * t_r4(tid) + i : -1;
        return i < NREG ? (i * NW + (tid >> 5)) * RPU + (tid & 31) / Q : -1;
    }
};

// The MMAs of one chunk: D[128 x (N0 | N1)] (+)= A[128 x KC] B[(N0 | N1) x KC]^T, `nterms` = 3: A_lo B_hi + A_hi B_lo +
// A_hi B_hi, 1: A_hi B_hi only.  `fresh` overwrites the accumulators (first chunk of a tile).
template <int KC>
__device__ __forceinline__ void issue_chunk(uint32_t tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int N0, int N1, int nterms, bool fresh)
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int nn = half == 0 ? N0 : N1;
        if (nn == 0) continue;
        const uint32_t idesc = instr_desc(128, nn);
        const uint32_t row0 = half == 0 ? 0u : (uint32_t)(N0 * row_bytes(KC));
        const uint32_t d = tmem + (half == 0 ? 0u : (uint32_t)N0);
        uint32_t acc = fresh ? 0u : 1u;
        for (int term = 3 - nterms; term < 3; ++term) {
            const uint32_t a = term == 0 ? a_lo : a_hi, b = (term == 1 ? b_lo : b_hi) + row0;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
                mma_tf32(d, smem_desc<KC>(a + ks * 32), smem_desc<KC>(b + ks * 32), idesc, acc);   // a k-step = 8 elements = 32 bytes into the row
                acc = 1u;
            }
        }
    }
}

struct SimStats {                       // double accumulators of the similarity-side loss terms (+ min, ordered-int key)
    double simval, ent, rec;
    int min_simval_key;
    int pad;
};
__device__ __forceinline__ int float_key_tc(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }

// ---------------------------------------------------------------------------------------------------------------------
// k_sim_tc: see the file header.  Per tile of 128 pixels:
//   chunks c = 0 .. nchunks-1 of the codebook width: [wait for the MMAs that read this stage] -> thread 0 starts the
//   bulk copies of the codebook images, all threads write the gt chunk (hi / lo, summing squares) and load the next
//   one into registers -> block barrier -> thread 0 waits for the bulk copies, issues the chunk's MMAs, commits.
//   epilogue: thread (row = t & 127, part = t >> 7) owns pixel `row` and the column range of its part; three passes
//   over its accumulator columns (tcgen05.ld, 32 at a time): max / arg-max -> exp sums -> gradient, label bits.
// Outputs: dsimT[tile][k][128] = d loss / d sim * (1 / |gt|) (zeros for padded rows / pixels), lmask[k / 32][pixel] =
// label bits (sim == row max), and the loss partial sums.
// ---------------------------------------------------------------------------------------------------------------------
template <int KC, bool PLANAR, int NT>                      // NT = 256 or 512 threads: NT / 128 column parts per pixel in the row pass
__global__ void __launch_bounds__(NT, 1)
k_sim_tc(int64_t N, int D, int K, int NP, int nchunks, int nterms, float t_anneal, const float* __restrict__ gt,
         const uint8_t* __restrict__ wimg, const int* __restrict__ zarg, float* __restrict__ dsimT,
         uint32_t* __restrict__ lmask, int64_t Npad, SimStats* __restrict__ stats, int ns)
{
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // swizzle atoms: 1024-byte aligned
    const int wbytes = NP * row_bytes(KC);                  // one codebook image (hi or lo) of one chunk
    constexpr int xbytes = 128 * row_bytes(KC);             // one gt image
    uint8_t* sW = smem_raw;                                 // [stage][hi, lo][wbytes]
    uint8_t* sX = sW + 2 * (size_t)ns * wbytes;             // [stage][hi, lo][xbytes]; ns = 2 .. MAX_STAGES stages
    __shared__ __align__(8) uint64_t s_wfull[MAX_STAGES], s_mma[MAX_STAGES], s_acc;
    __shared__ uint32_t s_tmem;
    __shared__ float s_ss[128];
    constexpr int NPARTS = NT / 128;
    using XTile = RowTile<KC, PLANAR, NT / 32>;
    __shared__ float s_pv[NPARTS][128], s_ps[NPARTS][128], s_pw[NPARTS][128];
    __shared__ int s_pi[NPARTS][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N0 = 16 * ((NP + 31) / 32), N1 = NP - N0;
    const bool want_lo = nterms == 3;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&s_wfull[s], 1); mbar_init(&s_mma[s], 1); }
        mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid < 128) s_ss[tid] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    const int64_t n_tiles = (N + 127) / 128;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(gt) & 15) == 0) && (((PLANAR ? N : (int64_t)D) & 3) == 0);
    const float invN = 1.0f / (float)N, ce = 0.3f * t_anneal * invN;
    double l_sim = 0.0, l_ent = 0.0, l_rec = 0.0;
    float l_min = INFINITY;

    // Two register sets: the loads of chunk g + 2 go out right after chunk g has been written to shared memory, so a
    // load has a whole chunk period to land before it is consumed.
    XTile xa, xb;
    auto load_x = [&](XTile& xt, int64_t it2, int c2) {
        // rows = pixels, reduction = codebook width; (it2, c2) may run past this tile: normalise
        it2 += c2 / nchunks;
        c2 %= nchunks;
        if (it2 < my_tiles) xt.load(gt, PLANAR ? N : (int64_t)D, (blockIdx.x + it2 * gridDim.x) * 128, N, (int64_t)c2 * KC, D, vec, tid);
    };
    load_x(xa, 0, 0);
    load_x(xb, 0, 1);

    uint32_t g = 0;                                         // chunks staged so far by this CTA
    int64_t tile = blockIdx.x;
    for (int64_t it = 0; it < my_tiles; ++it, tile += gridDim.x) {
        float ss[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = 0; c < nchunks; ++c, ++g) {
            const int s = (int)(g % (uint32_t)ns);
            const uint32_t u = g / (uint32_t)ns;
            if (g >= (uint32_t)ns) wait(smem_u32(&s_mma[s]), (u - 1) & 1);     // the MMAs that read this stage have completed
            const uint32_t w_hi = smem_u32(sW) + (uint32_t)(2 * s) * wbytes, w_lo = w_hi + wbytes;
            const uint32_t x_hi = smem_u32(sX) + (uint32_t)(2 * s) * xbytes, x_lo = x_hi + xbytes;
            if (tid == 0) {
                const uint32_t bar = smem_u32(&s_wfull[s]);
                mbar_expect_tx(bar, (uint32_t)(want_lo ? 2 * wbytes : wbytes));
                const uint8_t* src = wimg + (size_t)c * 2 * wbytes;
                bulk_g2s(w_hi, src, (uint32_t)wbytes, bar);
                if (want_lo) bulk_g2s(w_lo, src + wbytes, (uint32_t)wbytes, bar);
            }
            if ((g & 1) == 0) xa.store(x_hi, x_lo, want_lo, tid, ss); else xb.store(x_hi, x_lo, want_lo, tid, ss);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA
            __syncthreads();
            // behind the fence (a fence in front of the loads would wait for them to land)
            if ((g & 1) == 0) load_x(xa, it, c + 2); else load_x(xb, it, c + 2);
            if (tid == 0) {
                wait(smem_u32(&s_wfull[s]), u & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                issue_chunk<KC>(tmem, x_hi, x_lo, w_hi, w_lo, N0, N1, nterms, c == 0);
                commit(smem_u32(&s_mma[s]));
                if (c == nchunks - 1) commit(smem_u32(&s_acc));
            }
        }
        // ---- 1 / |gt| of the tile's pixels
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = XTile::row_of(i, tid);
            if (r >= 0) atomicAdd(&s_ss[r], ss[i]);
        }
        __syncthreads();
        const int row = tid & 127, part = tid >> 7;
        const int64_t n = tile * 128 + row;
        const bool rv = n < N;
        const float inv = rv ? 1.f / sqrtf(s_ss[row]) : 0.f;
        const int za = rv ? __ldg(zarg + n) : -1;
        wait(smem_u32(&s_acc), (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");

        // ---- epilogue: the similarity row of pixel `row`, columns [cbeg, cend).  Everything is evaluated on the raw
        // accumulators r_k (sim_k = r_k / |gt|, a positive scale): a2_k = c2 (r_k - r_max) = log2(e) t (sim_k - sim_max).
        const int NB = (NP + 31) / 32;                      // 32-column blocks; part p owns blocks [NB p / NPARTS, NB (p + 1) / NPARTS)
        const int cbeg = 32 * (NB * part / NPARTS), cend = min(NP, 32 * (NB * (part + 1) / NPARTS));
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
        // pass 1: row maximum, its first index, and the accumulator of column k^ (two chains: even / odd columns)
        float bv0 = -INFINITY, bv1 = -INFINITY, rec_raw = 0.f;
        int bi0 = 0, bi1 = 0;
        for_blocks(lane_base, cbeg, cend, K, [&](const uint32_t (&v)[32], int c0, int cnt, bool full) {
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float r0 = __uint_as_float(v[j]), r1 = __uint_as_float(v[j + 1]);
                    if (r0 > bv0) { bv0 = r0; bi0 = c0 + j; }
                    if (r1 > bv1) { bv1 = r1; bi1 = c0 + j + 1; }
                    rec_raw = (c0 + j == za) ? r0 : rec_raw;
                    rec_raw = (c0 + j + 1 == za) ? r1 : rec_raw;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float r = __uint_as_float(v[j]);
                    if (j < cnt && c0 + j < K) {
                        if (r > bv0) { bv0 = r; bi0 = c0 + j; }
                        rec_raw = (c0 + j == za) ? r : rec_raw;
                    }
                }
            }
        });
        if (bv1 > bv0 || (bv1 == bv0 && bi1 < bi0)) { bv0 = bv1; bi0 = bi1; }     // first maximum
        s_pv[part][row] = bv0;
        s_pi[part][row] = bi0;
        __syncthreads();
        float rmax = s_pv[0][row];
        int sarg = s_pi[0][row];
#pragma unroll
        for (int q = 1; q < NPARTS; ++q)                    // ascending parts, strict >: the first maximum wins
            if (s_pv[q][row] > rmax) { rmax = s_pv[q][row]; sarg = s_pi[q][row]; }
        const float c2 = LOG2E * t_anneal * inv;
        // pass 2: sum 2^a2 and sum 2^a2 a2
        float asum0 = 0.f, asum1 = 0.f, wsum0 = 0.f, wsum1 = 0.f;
        for_blocks(lane_base, cbeg, cend, K, [&](const uint32_t (&v)[32], int c0, int cnt, bool full) {
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float a0 = c2 * (__uint_as_float(v[j]) - rmax), a1 = c2 * (__uint_as_float(v[j + 1]) - rmax);
                    const float p0 = ex2_approx(a0), p1 = ex2_approx(a1);
                    asum0 += p0; asum1 += p1;
                    wsum0 = fmaf(p0, a0, wsum0); wsum1 = fmaf(p1, a1, wsum1);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float a0 = c2 * (__uint_as_float(v[j]) - rmax);
                    const float p0 = ex2_approx(a0);
                    if (j < cnt && c0 + j < K) { asum0 += p0; wsum0 = fmaf(p0, a0, wsum0); }
                }
            }
        });
        s_ps[part][row] = asum0 + asum1;
        s_pw[part][row] = wsum0 + wsum1;
        __syncthreads();
        float asum = 0.f, wsum = 0.f;
#pragma unroll
        for (int q = 0; q < NPARTS; ++q) { asum += s_ps[q][row]; wsum += s_pw[q][row]; }
        wsum *= LN2;
        const float ainv = 1.f / asum, logZ = __logf(asum);
        const float E = wsum * ainv - logZ;                 // sum P log P,  P = softmax(t sim)
        // pass 3: d loss / d sim_k * (1 / |gt|) = kd 2^a2 (ln2 a2 - (logZ + E)) (+ the two one-hot terms below), label bits
        const float kd = -ce * ainv * inv, LE = logZ + E;
        float* const out = dsimT + (size_t)tile * NP * 128 + row;
        for_blocks(lane_base, cbeg, cend, K, [&](const uint32_t (&v)[32], int c0, int cnt, bool full) {
            uint32_t bits = 0;
            if (full) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float r = __uint_as_float(v[j]);
                    const float a2 = c2 * (r - rmax);
                    out[(size_t)(c0 + j) * 128] = kd * ex2_approx(a2) * fmaf(a2, LN2, -LE);
                    bits |= (r == rmax) ? (1u << j) : 0u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < cnt) {
                        const float r = __uint_as_float(v[j]);
                        const float a2 = c2 * (r - rmax);
                        const bool real = c0 + j < K;
                        out[(size_t)(c0 + j) * 128] = real ? kd * ex2_approx(a2) * fmaf(a2, LN2, -LE) : 0.f;
                        bits |= (real && r == rmax) ? (1u << j) : 0u;
                    }
                }
            }
            lmask[(size_t)(c0 >> 5) * Npad + (size_t)n] = bits;
        });
        // the one-hot terms -(1/N)([k = k*] + [k = k^]) land on the columns this thread wrote itself
        const float oh = invN * inv;
        if (sarg >= cbeg && sarg < cend) out[(size_t)sarg * 128] -= oh;
        if (za >= cbeg && za < cend) out[(size_t)za * 128] -= oh;
        const float smax = rmax * inv, rec = rec_raw * inv;
        if (rv) {
            l_rec += (double)rec;
            if (part == 0) { l_sim += (double)smax; l_ent += (double)E; l_min = fminf(l_min, smax); }
        }
        if (tid < 128) s_ss[tid] = 0.f;
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();                                    // TMEM, s_ss and the exchange arrays are free for the next tile
    }

    // ---- loss partial sums
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        l_sim += __shfl_xor_sync(0xffffffffu, l_sim, o);
        l_ent += __shfl_xor_sync(0xffffffffu, l_ent, o);
        l_rec += __shfl_xor_sync(0xffffffffu, l_rec, o);
        l_min = fminf(l_min, __shfl_xor_sync(0xffffffffu, l_min, o));
    }
    if (lane == 0 && my_tiles > 0) {
        atomicAdd(&stats->simval, l_sim);
        atomicAdd(&stats->ent, l_ent);
        atomicAdd(&stats->rec, l_rec);
        if (l_min < INFINITY) atomicMin(&stats->min_simval_key, float_key_tc(l_min));
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
// k_dlut_tc: dlut1[k, d0 + r] += sum over this CTA's pixels of dsim[p, k] gt[p, d0 + r].  CTA = (slice of 128 codebook-
// width columns, every n_ranges-th pixel tile).  A operand = the gt slice (rows = d, reduction = pixels), B operand =
// dsimT (rows = codebook rows, reduction = pixels), accumulators [128 x NP] fp32 in TMEM for the whole kernel.
// ---------------------------------------------------------------------------------------------------------------------
template <int KC, bool PLANAR, int NT>                      // NT = 256 or 512 threads
__global__ void __launch_bounds__(NT, 1)
k_dlut_tc(int64_t N, int D, int K, int NP, int nterms, int n_slices, const float* __restrict__ gt,
          const float* __restrict__ dsimT, float* __restrict__ dlut1, int ns)
{
    constexpr int Q = KC / 4, RPU = 32 / Q;                 // pieces per row; rows per warp instruction
    constexpr int CPT = 128 / KC;                           // chunks per pixel tile
    constexpr int NW = NT / 32;                             // warps
    constexpr int BIT = ((KC == 32 ? 76 : 64) + NW - 1) / NW;      // B-operand pieces per thread: ceil(NP / RPU / NW), NP <= 304 | 512
    extern __shared__ uint8_t smem_dyn[];
    uint8_t* const smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int bbytes = NP * row_bytes(KC);
    constexpr int abytes = 128 * row_bytes(KC);
    uint8_t* sB = smem_raw;                                 // [stage][hi, lo][bbytes]
    uint8_t* sA = sB + 2 * (size_t)ns * bbytes;             // [stage][hi, lo][abytes]
    __shared__ __align__(8) uint64_t s_mma[MAX_STAGES], s_acc;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N0 = 16 * ((NP + 31) / 32), N1 = NP - N0;
    const bool want_lo = nterms == 3;
    const int slice = blockIdx.x % n_slices, range = blockIdx.x / n_slices, n_ranges = gridDim.x / n_slices;
    const int64_t d0 = (int64_t)slice * 128;

    if (tid == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) mbar_init(&s_mma[s], 1);
        mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    const int64_t n_tiles = (N + 127) / 128;
    const int64_t my_tiles = range < n_tiles ? (n_tiles - range + n_ranges - 1) / n_ranges : 0;
    const int64_t my_chunks = my_tiles * CPT;
    const bool vec = ((reinterpret_cast<uintptr_t>(gt) & 15) == 0) && (((PLANAR ? N : (int64_t)D) & 3) == 0);
    const int b_units = NP / RPU;

    // PLANAR: gt[d][pixel] is contiguous along the reduction (pixels) -> direct; row-major gt[pixel][d] -> transposed.
    // Two register sets, loads two chunks ahead (see k_sim_tc).
    struct Regs { RowTile<KC, !PLANAR, NW> at; float4 bv[BIT]; };
    Regs ra, rb;
    auto load_ab = [&](Regs& rg, int64_t ch) {
        if (ch >= my_chunks) return;
        const int64_t tile = range + (ch / CPT) * n_ranges;
        const int64_t p0 = tile * 128 + (ch % CPT) * KC;
        rg.at.load(gt, PLANAR ? N : (int64_t)D, d0, D, p0, N, vec, tid);
        const float* bsrc = dsimT + (size_t)tile * NP * 128 + (ch % CPT) * KC;
#pragma unroll
        for (int i = 0; i < BIT; ++i) {
            const int u = i * NW + warp;
            const int r = u * RPU + lane / Q, k4 = lane % Q;    // a warp instruction reads whole 128-byte (64-byte) rows
            if (u < b_units) rg.bv[i] = __ldg(reinterpret_cast<const float4*>(bsrc + (size_t)r * 128 + 4 * k4));
        }
    };
    load_ab(ra, 0);
    load_ab(rb, 1);

    float ss_unused[4] = {0.f, 0.f, 0.f, 0.f};
    auto step = [&](Regs& rg, int64_t ch) {
        const int s = (int)(ch % ns);
        const uint32_t u = (uint32_t)(ch / ns);
        if (ch >= ns) wait(smem_u32(&s_mma[s]), (u - 1) & 1);
        const uint32_t b_hi = smem_u32(sB) + (uint32_t)(2 * s) * bbytes, b_lo = b_hi + bbytes;
        const uint32_t a_hi = smem_u32(sA) + (uint32_t)(2 * s) * abytes, a_lo = a_hi + abytes;
        rg.at.store(a_hi, a_lo, want_lo, tid, ss_unused);
#pragma unroll
        for (int i = 0; i < BIT; ++i) {
            const int uu = i * NW + warp;
            const int r = uu * RPU + lane / Q, k4 = lane % Q;
            if (uu < b_units)
                store_split(b_hi, b_lo, (uint32_t)piece_off(KC, r, k4), rg.bv[i].x, rg.bv[i].y, rg.bv[i].z, rg.bv[i].w, want_lo);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        load_ab(rg, ch + 2);                                // behind the fence: a whole chunk period to land
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
            issue_chunk<KC>(tmem, a_hi, a_lo, b_hi, b_lo, N0, N1, nterms, ch == 0);
            commit(smem_u32(&s_mma[s]));
            if (ch == my_chunks - 1) commit(smem_u32(&s_acc));
        }
    };
    for (int64_t ch = 0; ch < my_chunks; ch += 2) {
        step(ra, ch);
        if (ch + 1 < my_chunks) step(rb, ch + 1);
    }

    if (my_chunks > 0) {
        wait(smem_u32(&s_acc), 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        constexpr int NPARTS = NT / 128;
        const int row = tid & 127, part = tid >> 7;
        const int NB = (NP + 31) / 32;
        const int cbeg = 32 * (NB * part / NPARTS), cend = min(NP, 32 * (NB * (part + 1) / NPARTS));
        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const bool dv = d0 + row < D;
        float* const out = dlut1 + d0 + row;
        uint32_t v[32];
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            const int cnt = min(32, cend - c0);
            ld_cols(lane_base + (uint32_t)c0, cnt, v);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < cnt && c0 + j < K && dv) atomicAdd(out + (size_t)(c0 + j) * D, __uint_as_float(v[j]));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
// k_zarg_tc: k^ = argmax_k (W x + b)_k per pixel (first maximum, like torch.argmax) -- the codebook row train.py:156
// looks up.  The S -> K projection of 128 pixels is one batch of tcgen05 MMAs (3 x TF32, the bias folded in as one
// more reduction column: x carries a 1 there, padded codebook rows a -3e38), the arg-max is one sweep over the
// accumulator row out of tensor memory; thread (row = t & 127, part = t >> 7) owns a pixel and half of the columns,
// ascending with a strict > so that the first maximum wins, part 0 winning ties against part 1.
// Operands (small: KP = S + 1 rounded up to 8 reduction elements) use the no-swizzle K-major layout of the mask kernel:
// element (row r, k) at (r / 8) SBO + (k / 4) 128 + (r % 8) 16 + (k % 4) 4, SBO = (KP / 4) 128.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc_plain(uint32_t saddr, int sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address
    d |= (uint64_t)(128 >> 4) << 16;                        // leading byte offset: the two 16-byte K pieces of a k-step
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;             // stride byte offset (8-row groups)
    d |= (uint64_t)1 << 46;                                 // descriptor version 1 (sm_100)
    return d;                                               // SWIZZLE_NONE
}

template <int NPARTS>                                       // column parts per pixel = warps per TMEM lane quarter (2 or 4)
__global__ void __launch_bounds__(128 * NPARTS, 1)
k_zarg_tc(int64_t N, int S, int K, int NP, int KP, const float* __restrict__ x, int64_t xs_n, int64_t xs_c,
          const float* __restrict__ W, const float* __restrict__ bias, int* __restrict__ zarg)
{
    constexpr int NT = 128 * NPARTS;
    extern __shared__ __align__(128) uint8_t smem_z[];
    const int SBO = (KP / 4) * 128;
    const int w_bytes = (NP / 8) * SBO, x_bytes = 16 * SBO;
    uint8_t* sWhi = smem_z;
    uint8_t* sWlo = sWhi + w_bytes;
    uint8_t* sXhi = sWlo + w_bytes;
    uint8_t* sXlo = sXhi + x_bytes;
    __shared__ __align__(8) uint64_t s_acc;
    __shared__ uint32_t s_tmem;
    __shared__ float s_pv[NPARTS][128];
    __shared__ int s_pi[NPARTS][128];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & 127, part = tid >> 7;
    const int N0 = 16 * ((NP + 31) / 32), N1 = NP - N0;
    auto elem_off = [SBO](int r, int k) { return (r >> 3) * SBO + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; };

    for (int i = tid; i < NP * KP; i += NT) {                // projection + bias column; padded rows can never win
        const int r = i / KP, k = i - r * KP;
        float v = 0.f;
        if (r < K) v = k < S ? W[(size_t)r * S + k] : (k == S ? (bias ? bias[r] : 0.f) : 0.f);
        else if (k == S) v = -3.0e38f;
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        *reinterpret_cast<uint32_t*>(sWhi + elem_off(r, k)) = hi;
        *reinterpret_cast<float*>(sWlo + elem_off(r, k)) = v - __uint_as_float(hi);
    }
    if (tid == 0) {
        mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    const int64_t n_tiles = (N + 127) / 128;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    constexpr int NPC = (10 + NPARTS - 1) / NPARTS;          // KP <= 40: ten 4-channel pieces; a thread stages pieces part, part + NPARTS, ..
    float xv[4 * NPC];
    auto load_x = [&](int64_t tile) {
        const int64_t n = tile * 128 + row;
        const float* px = x + n * xs_n;
#pragma unroll
        for (int q = 0; q < NPC; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 4 * (part + NPARTS * q) + j;
                xv[4 * q + j] = (k < S && n < N) ? __ldg(px + k * xs_c) : (k == S ? 1.f : 0.f);
            }
    };
    if (my_tiles > 0) load_x(blockIdx.x);
    const int NB = (NP + 31) / 32;                          // 32-column blocks; part p sweeps blocks [NB p / NPARTS, NB (p + 1) / NPARTS)
    const int cbeg = 32 * (NB * part / NPARTS), cend = min(NP, 32 * (NB * (part + 1) / NPARTS));
    int64_t tile = blockIdx.x;
    for (int64_t it = 0; it < my_tiles; ++it, tile += gridDim.x) {
        // operands of this tile (the previous tile's MMAs have completed: every thread waited for them below)
#pragma unroll
        for (int q = 0; q < NPC; ++q) {
            const int k4 = 4 * (part + NPARTS * q);
            if (k4 < KP) store_split(smem_u32(sXhi), smem_u32(sXlo), (uint32_t)elem_off(row, k4), xv[4 * q], xv[4 * q + 1], xv[4 * q + 2], xv[4 * q + 3], true);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");  // (the previous tile's tcgen05.ld of every thread)
        __syncthreads();
        if (it + 1 < my_tiles) load_x(tile + gridDim.x);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int nn = half == 0 ? N0 : N1;
                if (nn == 0) continue;
                const uint32_t idesc = instr_desc(128, nn);
                const uint32_t row0 = half == 0 ? 0u : (uint32_t)((N0 / 8) * SBO);
                const uint32_t d = tmem + (half == 0 ? 0u : (uint32_t)N0);
                uint32_t acc = 0u;
                for (int term = 0; term < 3; ++term) {      // small terms first: X_lo W_hi, X_hi W_lo, X_hi W_hi
                    const uint32_t a = smem_u32(term == 0 ? sXlo : sXhi), b = smem_u32(term == 1 ? sWlo : sWhi) + row0;
                    for (int ks = 0; ks < KP / 8; ++ks) {
                        mma_tf32(d, smem_desc_plain(a + ks * 256, SBO), smem_desc_plain(b + ks * 256, SBO), idesc, acc);
                        acc = 1u;
                    }
                }
            }
            commit(smem_u32(&s_acc));
        }
        wait(smem_u32(&s_acc), (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");

        const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        float bv0 = -INFINITY, bv1 = -INFINITY;
        int bi0 = 0, bi1 = 0;
        uint32_t v[32];
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            const int cnt = min(32, cend - c0);
            ld_cols(lane_base + (uint32_t)c0, cnt, v);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                if (j < cnt) {                               // cnt is 16 or 32: whole pairs
                    const float r0 = __uint_as_float(v[j]), r1 = __uint_as_float(v[j + 1]);
                    if (r0 > bv0) { bv0 = r0; bi0 = c0 + j; }
                    if (r1 > bv1) { bv1 = r1; bi1 = c0 + j + 1; }
                }
            }
        }
        if (bv1 > bv0 || (bv1 == bv0 && bi1 < bi0)) { bv0 = bv1; bi0 = bi1; }     // first maximum
        s_pv[part][row] = bv0;
        s_pi[part][row] = bi0;
        __syncthreads();
        const int64_t n = tile * 128 + row;
        if (part == 0 && n < N) {
#pragma unroll
            for (int q = 1; q < NPARTS; ++q)                 // ascending parts, strict >: the first maximum wins
                if (s_pv[q][row] > bv0) { bv0 = s_pv[q][row]; bi0 = s_pi[q][row]; }
            zarg[n] = bi0;
        }
        // (s_pv / s_pi are rewritten only after the next tile's block barrier)
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
// k_logit_tc: the logit side of the loss for S <= 16 channels and K <= 320 codebook rows (train.py:143-144, 154 and
// their backward) with every contraction on tensor cores:
//   z = W x + b for 128 pixels          tcgen05 MMAs into TMEM (as k_zarg_tc)
//   row pass out of TMEM, thread (row = t & 127, part = t >> 7) = (pixel, half of the 32-column blocks):
//     sweep 1  row maximum;  sweep 2  e = exp(z - max): sum e, sum e^2, sum of e over the label bits -- lab = sum (P' - L)^2
//     and dot = sum P' g follow from these three sums;  sweep 3  dz = P' (cl (P' - L) - dot), one 32-column block per
//     part at a time into a shared-memory block buffer [pixel][32 codebook rows]
//   per block pair, all eight warps (warp-level mma.sync m16n8k8, TF32 hi + lo split = fp32-accurate):
//     dx[16 pixels of the warp x S] += dz_block W_block          (accumulators in registers for the whole tile)
//     dW_block[32 rows x S] (+ db through a ones column) = dz_block^T x over a quarter of the pixels, flushed into a
//     shared-memory dW / db accumulator with red.shared
// The FMA kernel it replaces (k_semloss_rows, still used for S > 16 or K > 320) was bound by dependent LDS -> FMA chains.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_sync_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo)
{
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

template <int NTN>                                          // 8-channel tiles: 1 (S <= 8) or 2 (S <= 16)
__global__ void __launch_bounds__(THREADS, 1)
k_logit_tc(int64_t N, int S, int K, int NP, int KP, const float* __restrict__ x, int64_t xs_n, int64_t xs_c,
           const uint32_t* __restrict__ lmask, int64_t Npad, const float* __restrict__ W,
           const float* __restrict__ bias, float* __restrict__ dL_dx, float* __restrict__ dW, float* __restrict__ db,
           double* __restrict__ lab_out)
{
    constexpr int SPS = 8 * NTN + 8;                        // row stride (floats) of s_x, s_wf and s_dw: 16 or 24 -> conflict-free B fragments
    constexpr int DS = 36;                                  // row stride of a dz block: conflict-free A fragments of dx, 16-byte aligned rows
    constexpr int KPAD = 320;                               // codebook rows padded to ten 32-row blocks
    extern __shared__ __align__(128) uint8_t smem_l[];
    const int SBO = (KP / 4) * 128;
    const int w_bytes = (NP / 8) * SBO, x_bytes = 16 * SBO;
    uint8_t* sWhi = smem_l;
    uint8_t* sWlo = sWhi + w_bytes;
    uint8_t* sXhi = sWlo + w_bytes;
    uint8_t* sXlo = sXhi + x_bytes;
    uint32_t* s_wfh = reinterpret_cast<uint32_t*>(sXlo + x_bytes);   // [KPAD][SPS]   W[k][c] (zeros outside K x S), TF32 hi bits: B operand of dx
    uint32_t* s_wfl = s_wfh + KPAD * SPS;                            // [KPAD][SPS]   ... lo parts
    float* s_x = reinterpret_cast<float*>(s_wfl + KPAD * SPS);       // [128][SPS]    features of the tile: B operand of dW
    float* s_dz = s_x + 128 * SPS;                             // [2][128][DS]  one dz block per part
    float* s_dw = s_dz + 2 * 128 * DS;                         // [KPAD][SPS]   dW (columns < S) and db (column 8 NTN) of this CTA
    __shared__ __align__(8) uint64_t s_acc;
    __shared__ uint32_t s_tmem;
    __shared__ float s_e0[2][128], s_e1[2][128], s_e2[2][128], s_e4[2][128];
    __shared__ int s_e3[2][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
    const int row = tid & 127, part = tid >> 7;
    const int N0 = 16 * ((NP + 31) / 32), N1 = NP - N0;
    const int NB = (NP + 31) / 32, NB0 = (NB + 1) / 2;      // 32-column blocks; part 0 owns [0, NB0), part 1 [NB0, NB)
    auto elem_off = [SBO](int r, int k) { return (r >> 3) * SBO + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4; };

    for (int i = tid; i < NP * KP; i += THREADS) {           // logits operand: projection + bias column; padded rows -> -3e38
        const int r = i / KP, k = i - r * KP;
        float v = 0.f;
        if (r < K) v = k < S ? W[(size_t)r * S + k] : (k == S ? (bias ? bias[r] : 0.f) : 0.f);
        else if (k == S) v = -3.0e38f;
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        *reinterpret_cast<uint32_t*>(sWhi + elem_off(r, k)) = hi;
        *reinterpret_cast<float*>(sWlo + elem_off(r, k)) = v - __uint_as_float(hi);
    }
    for (int i = tid; i < KPAD * SPS; i += THREADS) {
        const int k = i / SPS, c = i - k * SPS;
        split_tf32((k < K && c < S) ? W[(size_t)k * S + c] : 0.f, s_wfh[i], s_wfl[i]);
        s_dw[i] = 0.f;
    }
    for (int i = tid; i < 128 * SPS; i += THREADS) s_x[i] = 0.f;
    if (tid == 0) {
        mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    const int64_t n_tiles = (N + 127) / 128;
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const float cl = 100.0f / ((float)N * (float)K);        // d(50 * MSE)/d(P') = 2 * 50 / (N K) * (P' - L)
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr int HALF_K = 20;                               // KP <= 24 here; part 0 stages channels 0 .. 19, part 1 the rest
    float xv[HALF_K];
    auto load_x = [&](int64_t tile) {
        const int64_t n = tile * 128 + row;
        const float* px = x + n * xs_n;
#pragma unroll
        for (int j = 0; j < HALF_K; ++j) {
            const int k = HALF_K * part + j;
            xv[j] = (k < S && n < N) ? __ldg(px + k * xs_c) : (k == S ? 1.f : 0.f);
        }
    };
    if (my_tiles > 0) load_x(blockIdx.x);
    double l_lab = 0.0;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16);

    int64_t tile = blockIdx.x;
    for (int64_t it = 0; it < my_tiles; ++it, tile += gridDim.x) {
        const int64_t n = tile * 128 + row;
        const bool rv = n < N;
        // ---- operands of this tile: TF32 images for the logits, plain fp32 rows for the dW product
#pragma unroll
        for (int j4 = 0; j4 < HALF_K; j4 += 4) {
            const int k4 = HALF_K * part + j4;
            if (k4 < KP) store_split(smem_u32(sXhi), smem_u32(sXlo), (uint32_t)elem_off(row, k4), xv[j4], xv[j4 + 1], xv[j4 + 2], xv[j4 + 3], true);
        }
        if (part == 0) {
#pragma unroll
            for (int c = 0; c < 8 * NTN; ++c) s_x[row * SPS + c] = c < S ? xv[c] : 0.f;
        }
        // label bits of this thread's blocks (bit j of word i = column 32 (first block + i) + j)
        uint32_t lw[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int b = (part ? NB0 : 0) + i;
            lw[i] = (rv && i < NB0 && b < NB) ? __ldg(lmask + (size_t)b * Npad + (size_t)n) : 0u;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (it + 1 < my_tiles) load_x(tile + gridDim.x);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int nn = half == 0 ? N0 : N1;
                if (nn == 0) continue;
                const uint32_t idesc = instr_desc(128, nn);
                const uint32_t row0 = half == 0 ? 0u : (uint32_t)((N0 / 8) * SBO);
                const uint32_t d = tmem + (half == 0 ? 0u : (uint32_t)N0);
                uint32_t acc = 0u;
                for (int term = 0; term < 3; ++term) {
                    const uint32_t a = smem_u32(term == 0 ? sXlo : sXhi), b = smem_u32(term == 1 ? sWlo : sWhi) + row0;
                    for (int ks = 0; ks < KP / 8; ++ks) {
                        mma_tf32(d, smem_desc_plain(a + ks * 256, SBO), smem_desc_plain(b + ks * 256, SBO), idesc, acc);
                        acc = 1u;
                    }
                }
            }
            commit(smem_u32(&s_acc));
        }
        wait(smem_u32(&s_acc), (uint32_t)(it & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");

        const int bfirst = part ? NB0 : 0, bend = part ? NB : NB0;
        uint32_t v[32];
        // ---- sweep 1: row maximum
        float zmax = -INFINITY;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int b = bfirst + i;
            if (b < bend) {
                const int c0 = 32 * b, cnt = min(32, NP - c0);
                ld_cols(lane_base + (uint32_t)c0, cnt, v);
                if (cnt == 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) zmax = fmaxf(zmax, __uint_as_float(v[j]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) zmax = fmaxf(zmax, __uint_as_float(v[j]));
                }
            }
        }
        s_e0[part][row] = zmax;
        __syncthreads();
        zmax = fmaxf(s_e0[0][row], s_e0[1][row]);
        const float zoff = -zmax * LOG2E;
        // ---- sweep 2: e = exp(z - max): sum e, sum e^2, sum over the label bits, number of label bits
        float es = 0.f, e2 = 0.f, eL = 0.f;
        int nL = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int b = bfirst + i;
            if (b < bend) {
                const int c0 = 32 * b, cnt = min(32, NP - c0);
                ld_cols(lane_base + (uint32_t)c0, cnt, v);
                const uint32_t bits = lw[i];
                nL += __popc(bits);
                if (cnt < 32) {                             // a 16-column last block: the upper half reads as "very negative"
#pragma unroll
                    for (int j = 16; j < 32; ++j) v[j] = 0xff000000u;      // -1.7e38
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float e = ex2_approx(fmaf(__uint_as_float(v[j]), LOG2E, zoff));      // padded rows: 2^-huge = 0
                    es += e;
                    e2 = fmaf(e, e, e2);
                    eL += ((bits >> j) & 1u) ? e : 0.f;
                }
            }
        }
        s_e1[part][row] = es; s_e2[part][row] = e2; s_e4[part][row] = eL; s_e3[part][row] = nL;
        __syncthreads();
        es = s_e1[0][row] + s_e1[1][row];
        e2 = s_e2[0][row] + s_e2[1][row];
        eL = s_e4[0][row] + s_e4[1][row];
        nL = s_e3[0][row] + s_e3[1][row];
        const float zinv = 1.f / es;
        const float sP2 = e2 * zinv * zinv, sPL = eL * zinv;      // sum P'^2, sum of P' over the label bits
        const float dot = cl * (sP2 - sPL);                       // sum P' g, g = cl (P' - L)
        if (rv && part == 0) l_lab += (double)(sP2 - 2.f * sPL + (float)nL);      // sum (P' - L)^2

        // ---- sweep 3 + the two products, one 32-column block per part at a time
        float accx[NTN][4];
#pragma unroll
        for (int nt = 0; nt < NTN; ++nt) { accx[nt][0] = accx[nt][1] = accx[nt][2] = accx[nt][3] = 0.f; }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i >= NB0) break;
            const int b = bfirst + i;
            float* dzrow = s_dz + ((size_t)part * 128 + row) * DS;
            if (b < bend) {
                const int c0 = 32 * b, cnt = min(32, NP - c0);
                ld_cols(lane_base + (uint32_t)c0, cnt, v);
                const uint32_t bits = lw[i];
                if (cnt < 32) {
#pragma unroll
                    for (int j = 16; j < 32; ++j) v[j] = 0xff000000u;      // -1.7e38 -> P' = 0 -> dz = 0
                }
                const float zi = rv ? zinv : 0.f;           // pixels past the end: P' = 0 -> dz = 0
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float d4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float Pz = ex2_approx(fmaf(__uint_as_float(v[j + q]), LOG2E, zoff)) * zi;
                        d4[q] = Pz * fmaf(cl, Pz - (((bits >> (j + q)) & 1u) ? 1.f : 0.f), -dot);      // padded rows: P' = 0
                    }
                    *reinterpret_cast<float4*>(dzrow + j) = make_float4(d4[0], d4[1], d4[2], d4[3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dzrow + j) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {
                const int bb = (blk ? NB0 : 0) + i;          // global block index of this buffer
                if (bb >= (blk ? NB : NB0)) continue;
                const float* dzb = s_dz + (size_t)blk * 128 * DS;
                // dx: this warp's 16 pixels x the block's 32 codebook rows (4 k-steps)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const float* ap = dzb + (16 * warp + gid) * DS + 8 * ks + tig;
                    uint32_t ah[4], al[4];
                    split_tf32(ap[0], ah[0], al[0]); split_tf32(ap[8 * DS], ah[1], al[1]);
                    split_tf32(ap[4], ah[2], al[2]); split_tf32(ap[8 * DS + 4], ah[3], al[3]);
#pragma unroll
                    for (int nt = 0; nt < NTN; ++nt) {
                        const int bo = (32 * bb + 8 * ks + tig) * SPS + 8 * nt + gid;
                        const uint32_t bh0 = s_wfh[bo], bh1 = s_wfh[bo + 4 * SPS], bl0 = s_wfl[bo], bl1 = s_wfl[bo + 4 * SPS];
                        mma_sync_tf32(accx[nt], al, bh0, bh1);
                        mma_sync_tf32(accx[nt], ah, bl0, bl1);
                        mma_sync_tf32(accx[nt], ah, bh0, bh1);
                    }
                }
                // dW / db: 16 of the block's rows (mt) x a quarter of the pixels (kq): 4 k-steps, then flush
                {
                    const int mt = warp & 1, kq = warp >> 1;
                    float accd[NTN + 1][4];
#pragma unroll
                    for (int nt = 0; nt <= NTN; ++nt) { accd[nt][0] = accd[nt][1] = accd[nt][2] = accd[nt][3] = 0.f; }
                    const uint32_t one = gid == 0 ? 0x3f800000u : 0u;      // ones tile: column 0 = 1 -> db
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const int px0 = 32 * kq + 8 * ks;
                        const float* ap = dzb + (px0 + tig) * DS + 16 * mt + gid;      // A[row = codebook row][k = pixel]
                        uint32_t ah[4], al[4];
                        split_tf32(ap[0], ah[0], al[0]); split_tf32(ap[8], ah[1], al[1]);
                        split_tf32(ap[4 * DS], ah[2], al[2]); split_tf32(ap[4 * DS + 8], ah[3], al[3]);
#pragma unroll
                        for (int nt = 0; nt < NTN; ++nt) {
                            const float* bp = s_x + (px0 + tig) * SPS + 8 * nt + gid;
                            uint32_t bh0, bl0, bh1, bl1;
                            split_tf32(bp[0], bh0, bl0); split_tf32(bp[4 * SPS], bh1, bl1);
                            mma_sync_tf32(accd[nt], al, bh0, bh1);
                            mma_sync_tf32(accd[nt], ah, bl0, bl1);
                            mma_sync_tf32(accd[nt], ah, bh0, bh1);
                        }
                        mma_sync_tf32(accd[NTN], al, one, one);
                        mma_sync_tf32(accd[NTN], ah, one, one);
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float* drow = s_dw + (32 * bb + 16 * mt + gid + 8 * (e >> 1)) * SPS;
#pragma unroll
                        for (int nt = 0; nt < NTN; ++nt) atomicAdd(drow + 8 * nt + 2 * tig + (e & 1), accd[nt][e]);
                        if (tig == 0 && (e & 1) == 0) atomicAdd(drow + 8 * NTN, accd[NTN][e]);
                    }
                }
            }
            __syncthreads();                                // the block buffers are rewritten by the next iteration
        }
        // ---- dx of the tile: C fragment rows = pixels 16 warp + gid (+ 8), columns = channels 8 nt + 2 tig (+ 1)
        if (dL_dx) {
#pragma unroll
            for (int nt = 0; nt < NTN; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t p = tile * 128 + 16 * warp + gid + 8 * (e >> 1);
                    const int c = 8 * nt + 2 * tig + (e & 1);
                    if (p < N && c < S) dL_dx[p * xs_n + c * xs_c] = accx[nt][e];
                }
        }
    }

    // ---- this CTA's dW / db and loss partial sum
    __syncthreads();
    if (dW) {
        for (int i = tid; i < K * (S + 1); i += THREADS) {
            const int k = i / (S + 1), c = i - k * (S + 1);
            if (c < S) atomicAdd(dW + (size_t)k * S + c, s_dw[k * SPS + c]);
            else if (db) atomicAdd(db + k, s_dw[k * SPS + 8 * NTN]);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) l_lab += __shfl_xor_sync(0xffffffffu, l_lab, o);
    if (lane == 0 && my_tiles > 0) atomicAdd(lab_out, l_lab);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// The codebook operand of k_sim_tc: lut1 split once into TF32 hi / lo images, chunk by chunk, already in the shared-
// memory layout (rows >= K and columns >= D are zeros) so that a stage is one bulk copy per image.
__global__ void __launch_bounds__(256) k_build_wimg(int K, int D, int NP, int KC, int nchunks,
                                                    const float* __restrict__ lut1, uint8_t* __restrict__ img)
{
    const int wbytes = NP * row_bytes(KC);
    const int total = nchunks * NP * KC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (NP * KC), rk = i % (NP * KC), r = rk / KC, k = rk % KC;
        const int d = c * KC + k;
        const float v = (r < K && d < D) ? lut1[(size_t)r * D + d] : 0.f;
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        const size_t off = (size_t)c * 2 * wbytes + (size_t)(piece_off(KC, r, k >> 2) + (k & 3) * 4);
        *reinterpret_cast<uint32_t*>(img + off) = hi;
        *reinterpret_cast<float*>(img + off + wbytes) = v - __uint_as_float(hi);
    }
}

}  // namespace tc5
