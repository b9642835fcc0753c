// soundtouch.cu -- K6/K7: the SoundTouch 2.3.2 pipeline behind pitch_modifier / velocity_modifier
// (A9; audio-velocity.cpp:265-443 drives SoundTouch::putSamples / receiveSamples / flush).
//
// SoundTouch is FIFO driven, so its output is a pure function of the whole input; the streaming
// object becomes four whole-track kernels (batched over tracks of equal length):
//
//   tds_offsets  TDStretch::seekBestOverlapPositionFull -- the only sequential part.  Input
//                consumption (skipFract) does not depend on the chosen offsets, so the host lays
//                out every sequence's input position up front; what remains is the chain
//                offset_i = argmax_c w(c) * (corr(in[pos_i + c ..], mid_i) + 0.1) with mid_i cut
//                from the input at the previous offset.  One CTA per track walks that chain.  Per
//                sequence the search window is staged once into shared memory, de-interleaved
//                into the four SSE lanes' sample streams; a thread owns one (lane, candidate
//                class) pair and T consecutive candidates, keeps T correlation and T norm
//                accumulators in registers and slides one window of samples over them, so a
//                shared-memory word feeds 2T rounded multiply/adds.  The four lane sums of a
//                candidate are combined in the SSE build's order ((l0+l1)+l2)+l3 and the arg-max
//                is first-wins like the scalar loop -- the offset trace is bit-identical.
//   tds_assemble cross-fade + sequence copy, fully parallel once the offsets are known.
//   aa_fir       AAFilter, 64 taps, the SSE build's even/odd accumulator order (stereo) or the
//                double accumulator (mono).
//   cubic        InterpolateCubic::transpose*.  rate = pitch*rate is a product of two floats, so
//                fract += rate is exact in double and the read position of output i is the
//                closed form i*rate in 128-bit fixed point: no sequential pass.
//
// Order: rate <= 1: cubic -> aa_fir -> TDStretch; rate > 1: TDStretch -> aa_fir -> cubic
// (SoundTouch::putSamples).  flush() becomes "extend the input with 128-frame silent blocks
// until enough output exists, trim to the expected total".
// Compiled with -fmad=false: every multiply and add rounds separately, as in the reference build.
#include "nodey_common.cuh"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace nodey {

constexpr int kAaLen = 64;
// 256 threads, two CTAs per SM, measured best: 384 / 512-thread CTAs (tools/micro: libnodey_cuda_t384/512.so + tools/st_sweep.py)
// were 3-5 % slower at 256 tracks and no faster at 32; a three-block register ring for the sample window changed nothing
#ifndef NODEY_TDS_THREADS
#define NODEY_TDS_THREADS 256
#endif
constexpr int kTdsThreads = NODEY_TDS_THREADS;
#ifndef NODEY_TDS_RESIDENT
#define NODEY_TDS_RESIDENT 2
#endif      // two CTAs (two tracks, or two slices of one) share an SM and fill each other's staging / reduction gaps

// frames of a (possibly batched) stream with virtual silence: `prefix` silent frames in front
// (RateTransposer latency pre-fill) and silence after `n` real frames (flush blocks).
// Two addressings: one contiguous batch of interleaved tracks `stride` floats apart (the intermediates), or a
// per-track pointer table (the node's input read in place: interleaved FLT, or the two planes of stereo FLTP --
// extract_samples_interleaved, audio-velocity.cpp:150-232, is the identity on float samples, so no copy is made).
struct View {
    const float* p;
    long long n;        // real frames
    long long prefix;   // silent frames in front
    long long stride;   // floats between tracks
    int use_tab;        // 1: tracks are addressed through the kernel's TrackTab instead of p + track * stride
};

// per-track source pointers, passed by value in the kernel parameters (no upload, no lifetime to manage):
// p[2t] = interleaved frames or left plane, p[2t + 1] = right plane or null
constexpr int kMaxTabTracks = 256;
struct TrackTab { const float* p[2 * kMaxTabTracks]; };

struct Src { const float* a; const float* b; };     // b != nullptr: stereo planes (a = left, b = right)

__device__ __forceinline__ Src view_src(const View& v, const TrackTab& tt, long long track)
{
    if (v.use_tab) return Src{tt.p[2 * track], tt.p[2 * track + 1]};
    return Src{v.p + track * v.stride, nullptr};
}
__device__ __forceinline__ Src view_src(const View& v, long long track) { return Src{v.p + track * v.stride, nullptr}; }

template <int CH>
__device__ __forceinline__ float view_sample(const View& v, const Src& base, long long frame, int c)
{
    const long long f = frame - v.prefix;
    if (f < 0 || f >= v.n) return 0.f;
    if (CH == 2 && base.b) return (c ? base.b : base.a)[f];
    return base.a[f * CH + c];
}

template <int CH>
__device__ __forceinline__ float2 view_frame2(const View& v, const Src& base, long long frame)
{
    const long long f = frame - v.prefix;
    if (f < 0 || f >= v.n) return make_float2(0.f, 0.f);
    if (CH == 2) {
        if (base.b) return make_float2(base.a[f], base.b[f]);
        return *reinterpret_cast<const float2*>(base.a + f * 2);
    }
    return make_float2(base.a[f], 0.f);
}

// ---------------------------------------------------------------------------------------------
// TDStretch offsets
// ---------------------------------------------------------------------------------------------
struct TdsArgs {
    View in; TrackTab tt;
    const long long* pos;      // [nseq] input frame where sequence i starts
    int* offs;                 // [ntracks][offs_stride] offsets of sequences 1..nseq-1 at index i-1
    long long offs_stride;
    int nseq;
    int overlap, seek_window, seek_length;
    int Q;                     // lane steps = 4 * (CH*overlap/16)
    int qp;                    // words per lane row of the staged mid buffer (groups of KT steps, each padded to a multiple of 4)
    int tb_per;                // candidate blocks per CTA of the cluster
    int sk;                    // floats per sub-plane
    int ncand_pad;             // padded candidates per CTA (correlation lane sums)
    int npm;                   // padded start positions per CTA (norm sums)
    int vec8;                  // stereo frames are 8-byte aligned in global memory
    int seq_begin, seq_end;    // this launch searches sequences [seq_begin, seq_end) (1 <= seq_begin): a track's chain can be
                               // cut into several launches; the only carried state is the previous offset, read from offs[]
};

// partial-sum slot of candidate cc: one word of padding per K*KT candidates makes both the strided
// stores of the lane-sum phase (stride K*KT) and the unit-stride loads of the combine phase conflict free
// (two words when S is odd, so that the stride S + pad stays odd)
template <int S>
__device__ __forceinline__ int ps_slot(int cc) { return cc + (cc / S) * ((S & 1) ? 2 : 1); }


#ifdef NODEY_TDS_TIMING
// development build only (tools/micro/Makefile): clock64 deltas of the phases of one sequence, thread 0 of CTA 0
__device__ unsigned long long g_tds_phase[8];
#define TDS_T(k) do { if (timing_on) { const long long now_ = clock64(); tacc[k] += now_ - tprev; tprev = now_; } } while (0)
#else
#define TDS_T(k) do { } while (0)
#endif

// One group of NS lane steps for KT consecutive stream positions of one (lane, class) stream.
// Window = two blocks of KT samples (w[CUR] current, w[CUR^1] next); sample idx = s + k of the
// window feeds position k at step s.  Block b of a thread's stream lives at
// plane[k * sk + b]: consecutive threads -> consecutive words.
// MODE 0: correlation lane sums only (acc[k] += x * y); MODE 1: norm sums only (acc[k] += x * x).
// The norm of a candidate does not involve the mid buffer, and candidates of different (lane, class)
// pairs walk the same squared samples: it is computed once per sample stream (plane) and start
// position instead of once per (lane, class) -- see tds_offsets_kernel.
// PARTIAL: only the first `nsteps` (< KT) steps run -- the last group of a row when KT does not divide Q.
template <int KT, int CUR, int MODE, bool PARTIAL>
__device__ __forceinline__ void tds_group(float (&w)[2][KT], float (&acc)[KT], const float* __restrict__ xnext, int sk,
                                          const float* __restrict__ yq, int nsteps = KT)
{
    constexpr int NXT = CUR ^ 1;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        const float v = xnext[k * sk];
        w[NXT][k] = MODE == 0 ? v : __fmul_rn(v, v);
    }
    float y[(KT + 3) & ~3];
    if (MODE == 0) {
        if constexpr (KT == 2) {
            const float2 t = *reinterpret_cast<const float2*>(yq);       // rows in groups of two words (tds_ktp)
            y[0] = t.x; y[1] = t.y;
        } else {
#pragma unroll
            for (int s = 0; s < KT; s += 4) {
                const float4 t = *reinterpret_cast<const float4*>(yq + s);
                y[s] = t.x; y[s + 1] = t.y; y[s + 2] = t.z; y[s + 3] = t.w;
            }
        }
    }
#pragma unroll
    for (int s = 0; s < KT; s++) {
        if (PARTIAL && s >= nsteps) break;
#pragma unroll
        for (int k = 0; k < KT; k++) {
            const int idx = s + k;
            const float x = idx < KT ? w[CUR][idx] : w[NXT][idx - KT];
            acc[k] = __fadd_rn(acc[k], MODE == 0 ? __fmul_rn(x, y[s]) : x);
        }
    }
}

// rows of the mid buffer are stored in groups of KT steps padded to a multiple of four words, so that every group
// starts 16-byte aligned whatever KT is (KT = 2: groups of two words, read as 64-bit pairs)
template <int KT> __host__ __device__ constexpr int tds_ktp() { return KT == 2 ? 2 : (KT + 3) & ~3; }
static int tds_ktp_rt(int KT) { return KT == 2 ? 2 : (KT + 3) & ~3; }

// Small-batch variant for KT = 4 (four CTAs per track): with one or two warps per scheduler a group's window block and mid
// buffer words were used a few multiplies after their loads were issued and the warp sat out the shared-memory latency
// every 37 instructions.  Here both ride one group ahead in registers: a ring of THREE window blocks and two mid-buffer
// groups (24 registers in all); six groups per iteration bring the ring back to its start, so every index is static.
// WI = ring slot of the current block, YI = slot of the current mid-buffer group.
template <int KT, int MODE, int WI, int YI>
__device__ __forceinline__ void tds_group_ahead(float (&w)[3][KT], float (&y)[2][KT], float (&acc)[KT], const float* __restrict__ xahead, int sk,
                                                const float* __restrict__ ynext)
{
    constexpr int NXT = (WI + 1) % 3, AH = (WI + 2) % 3;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        const float v = xahead[k * sk];
        w[AH][k] = MODE == 0 ? v : __fmul_rn(v, v);
    }
    if (MODE == 0) {
#pragma unroll
        for (int s = 0; s < KT; s += 4) {
            const float4 t = *reinterpret_cast<const float4*>(ynext + s);
            y[YI ^ 1][s] = t.x; y[YI ^ 1][s + 1] = t.y; y[YI ^ 1][s + 2] = t.z; y[YI ^ 1][s + 3] = t.w;
        }
    }
#pragma unroll
    for (int s = 0; s < KT; s++) {
#pragma unroll
        for (int k = 0; k < KT; k++) {
            const int idx = s + k;
            const float x = idx < KT ? w[WI][idx] : w[NXT][idx - KT];
            acc[k] = __fadd_rn(acc[k], MODE == 0 ? __fmul_rn(x, y[YI][s]) : x);
        }
    }
}

// block G + 2 and mid-buffer group G + 1 are loaded while group G is summed: the staged window holds Q / KT + 3 blocks per
// thread and a mid-buffer row eight spare words, so the loads of the last groups stay inside the arrays (values unused)
template <int KT, int MODE, int SK>
__device__ __forceinline__ void tds_lane_sums_ahead(const float* __restrict__ xb, int sk_rt, const float* __restrict__ yp, int Q, float (&acc)[KT])
{
    static_assert(KT % 4 == 0, "mid-buffer groups are read as 128-bit words");
    const int sk = SK ? SK : sk_rt;
    float w[3][KT], y[2][KT];
#pragma unroll
    for (int k = 0; k < KT; k++) acc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        const float v0 = xb[k * sk], v1 = xb[k * sk + 1];
        w[0][k] = MODE == 0 ? v0 : __fmul_rn(v0, v0);
        w[1][k] = MODE == 0 ? v1 : __fmul_rn(v1, v1);
    }
    if (MODE == 0) {
#pragma unroll
        for (int s = 0; s < KT; s += 4) {
            const float4 t = *reinterpret_cast<const float4*>(yp + s);
            y[0][s] = t.x; y[0][s + 1] = t.y; y[0][s + 2] = t.z; y[0][s + 3] = t.w;
        }
    }
    const int NG = Q / KT;                          // KT divides Q (Q is a multiple of four)
    int G = 0;
#pragma unroll 1
    for (; G + 6 <= NG; G += 6) {
        tds_group_ahead<KT, MODE, 0, 0>(w, y, acc, xb + G + 2, sk, yp + (G + 1) * KT);
        tds_group_ahead<KT, MODE, 1, 1>(w, y, acc, xb + G + 3, sk, yp + (G + 2) * KT);
        tds_group_ahead<KT, MODE, 2, 0>(w, y, acc, xb + G + 4, sk, yp + (G + 3) * KT);
        tds_group_ahead<KT, MODE, 0, 1>(w, y, acc, xb + G + 5, sk, yp + (G + 4) * KT);
        tds_group_ahead<KT, MODE, 1, 0>(w, y, acc, xb + G + 6, sk, yp + (G + 5) * KT);
        tds_group_ahead<KT, MODE, 2, 1>(w, y, acc, xb + G + 7, sk, yp + (G + 6) * KT);
    }
    const int r = NG - G;
    if (r > 0) tds_group_ahead<KT, MODE, 0, 0>(w, y, acc, xb + G + 2, sk, yp + (G + 1) * KT);
    if (r > 1) tds_group_ahead<KT, MODE, 1, 1>(w, y, acc, xb + G + 3, sk, yp + (G + 2) * KT);
    if (r > 2) tds_group_ahead<KT, MODE, 2, 0>(w, y, acc, xb + G + 4, sk, yp + (G + 3) * KT);
    if (r > 3) tds_group_ahead<KT, MODE, 0, 1>(w, y, acc, xb + G + 5, sk, yp + (G + 4) * KT);
    if (r > 4) tds_group_ahead<KT, MODE, 1, 0>(w, y, acc, xb + G + 6, sk, yp + (G + 5) * KT);
}

// ROT = false: the loop body holds two groups (the window blocks swap roles, no moves); ROT = true: one group per
// iteration and KT register moves -- half the code.  A body of two groups of KT = 15 is 15 KB of instructions per variant,
// and the variants live in the SM's 32 KB instruction cache together: measured, the large-KT kernels lost more to
// instruction fetch than they saved in issued instructions until their bodies were halved.
template <int KT, int MODE, int SK, bool ROT>
__device__ __forceinline__ void tds_lane_sums(const float* __restrict__ xb, int sk_rt, const float* __restrict__ yp, int Q,
                                              float (&acc)[KT])
{
    const int sk = SK ? SK : sk_rt;                 // compile-time sub-plane stride: every window load is base + immediate
    constexpr int KTP = tds_ktp<KT>();
    float w[2][KT];
#pragma unroll
    for (int k = 0; k < KT; k++) acc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < KT; k++) {
        const float v = xb[k * sk];
        w[0][k] = MODE == 0 ? v : __fmul_rn(v, v);
    }
    const int NG = Q / KT, tail = Q - NG * KT;
    int G = 0;
    if constexpr (ROT) {
#pragma unroll 1
        for (; G < NG; G++) {
            tds_group<KT, 0, MODE, false>(w, acc, xb + G + 1, sk, yp + G * KTP);
#pragma unroll
            for (int k = 0; k < KT; k++) w[0][k] = w[1][k];
        }
        if (tail) tds_group<KT, 0, MODE, true>(w, acc, xb + G + 1, sk, yp + G * KTP, tail);
    } else {
        if constexpr (KT == 2) {
            // a group of two steps is eight operations: four pairs of groups per iteration keep the loop overhead small
#pragma unroll 1
            for (; G + 8 <= NG; G += 8) {
#pragma unroll
                for (int g = 0; g < 8; g += 2) {
                    tds_group<KT, 0, MODE, false>(w, acc, xb + G + g + 1, sk, yp + (G + g) * KTP);
                    tds_group<KT, 1, MODE, false>(w, acc, xb + G + g + 2, sk, yp + (G + g + 1) * KTP);
                }
            }
        }
        for (; G + 2 <= NG; G += 2) {
            tds_group<KT, 0, MODE, false>(w, acc, xb + G + 1, sk, yp + G * KTP);
            tds_group<KT, 1, MODE, false>(w, acc, xb + G + 2, sk, yp + (G + 1) * KTP);
        }
        if (G < NG) {
            tds_group<KT, 0, MODE, false>(w, acc, xb + G + 1, sk, yp + G * KTP);
            G++;
            if (tail) tds_group<KT, 1, MODE, true>(w, acc, xb + G + 1, sk, yp + G * KTP, tail);
        } else if (tail) {
            tds_group<KT, 0, MODE, true>(w, acc, xb + G + 1, sk, yp + G * KTP, tail);
        }
    }
}

// ---- asynchronous global -> shared copies (LDGSTS): no registers held while the data is in flight ----
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = valid ? 4 : 0;                 // src-size 0: nothing is read, the word is zero filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async8(float* dst_smem, const float* src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- cluster exchange without a cluster barrier: remote stores that signal the receiver's mbarrier (st.async) ----
__device__ __forceinline__ unsigned st_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned st_map_rank(const void* smem_ptr, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(st_smem_u32(smem_ptr)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_mbar_init(unsigned long long* m, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(st_smem_u32(m)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive_expect(unsigned long long* m, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(st_smem_u32(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_async_pair(unsigned dst_cluster, unsigned long long v0, unsigned long long v1, unsigned mbar_cluster)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 :: "r"(dst_cluster), "l"(v0), "l"(v1), "r"(mbar_cluster) : "memory");
}
__device__ __forceinline__ void st_mbar_wait_cluster(unsigned long long* m, unsigned parity)
{
    unsigned done = 0, spins = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(st_smem_u32(m)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) __trap();      // a lost signal must not hang the device
    }
}

// order-preserving map double -> u64 for the arg-max: NaN lowest (the scalar loop's `>` never picks it), -0 == +0
__device__ __forceinline__ unsigned long long argmax_key(double v)
{
    if (!(v == v)) return 0ull;
    if (v == 0.0) v = 0.0;
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

// grid = ntracks * CL CTAs, launched as clusters of CL: the CTAs of a cluster split one track's
// candidates (blocks of KT per class) and exchange their local arg-max through distributed
// shared memory, one cluster barrier per sequence.
//
// What is on the per-sequence critical path and what is not (measured with the clock64 phase timers of the
// NODEY_TDS_TIMING build, tools/tds_phases.py): only the correlation lane sums need the mid buffer, i.e.
// the previous offset.  Everything else is moved off the chain:
//   * the search window of sequence i+1 and the REGION the next mid buffer can come from (it starts at
//     pos_i + overlap + temp + offset_i, offset_i < seek_length) are copied global -> shared with cp.async
//     while sequence i is searched; the mid buffer is then a shared -> shared gather;
//   * the norm sums of sequence i+1 (no mid buffer involved) are computed between the arrive and the wait
//     of the cluster barrier that publishes offset i;
//   * the position weights 1 - 0.25 t^2 are tabulated once.
// SK > 0: the sub-plane stride is a compile-time constant (a.sk == SK), 0: run-time a.sk
template <int CH, int KT, int SK>
__global__ void __launch_bounds__(kTdsThreads, KT == 2 ? 4 : NODEY_TDS_RESIDENT) tds_offsets_kernel(const __grid_constant__ TdsArgs a)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks(), crank = cluster.block_rank();

    extern __shared__ __align__(16) float smem[];
    constexpr int K = 4 / CH;                       // candidate classes per lane
    const int L = a.seek_length, ovl = a.overlap, Q = a.Q;
    constexpr int KTP = tds_ktp<KT>();
    const int QP = a.qp;                            // padded lane row ((Q + KT - 1) / KT groups of KTP words + 8): the de-interleaving stores hit 32 banks
    const int sk = SK ? SK : a.sk;
    const int plane_len = KT * sk;
    const int mreg = L + ovl;                       // frames the next mid buffer can come from
    float* X0 = smem;                               // 2 x [4 planes][KT sub-planes][sk] (double buffered)
    float* Y = X0 + 8 * plane_len;                  // [4][QP]
    float* PS = Y + 4 * QP;                         // [4][ncand_pad] correlation lane sums
    float* PN = PS + 4 * a.ncand_pad;               // [4 planes][npm] norm sums by start position
    float* MR = PN + 4 * a.npm;                     // [mreg * CH] interleaved frames: mid-buffer region
    double* PW = reinterpret_cast<double*>(MR + ((mreg * CH + 3) & ~3));   // [ncand] position weights
    int* SLT = reinterpret_cast<int*>(PW + K * KT * a.tb_per);               // [ncand][5] partial-sum slots of a candidate (see the combine phase)
    __shared__ unsigned long long red_k[kTdsThreads / 32];
    __shared__ int red_i[kTdsThreads / 32];
    __shared__ __align__(16) unsigned long long xch[2][8][2];     // [parity][sender rank]{key, index}: written by the peers (st.async)
    __shared__ __align__(8) unsigned long long xbar[2];           // one mbarrier per parity: counts the bytes of the CL results

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const long long track = blockIdx.x / CL;
    const Src base = view_src(a.in, a.tt, track);
    int* offs = a.offs + track * a.offs_stride;

    const int region = L + ovl;                      // frames of the search window
    const int temp = a.seek_window - 2 * ovl;
    const int tcount = (L + K - 1) / K;
    const int tblocks = (tcount + KT - 1) / KT;
    const int tb_lo = (int)crank * a.tb_per;
    int ntb = tblocks - tb_lo; if (ntb > a.tb_per) ntb = a.tb_per; if (ntb < 0) ntb = 0;
    constexpr bool ROT = KT > 8;                     // large bodies: one group per loop iteration (instruction cache)
    // four CTAs per track (few tracks per GPU): loads ride one group ahead (tds_lane_sums_ahead).  Measured in the engine: 16 tracks
    // 22.95 -> 20.16 ms, 32 tracks 31.1 -> 30.2 ms; the same ring for KT = 8 (two CTAs per track) made 64 tracks 2 % slower (50.9 -> 52.0 ms)
    constexpr bool kAhead = KT == 4;
    const int wpc = (ntb + 1 + 31) / 32;             // warps per (lane, class) combo (streams that start one element late need one more block)
    const int nunits = 4 * K * wpc;
    const int wpn = (ntb + 1 + 31) / 32;             // warps per plane for the norm sums (ntb + 1 position blocks)
    const int nnorm = 4 * wpn;
    const int m_lo = KT * tb_lo;                     // first plane element this CTA stages
    const int f_lo = 4 * m_lo / CH;                  // ... and the first frame
    // frames to stage: what the lane sums can read (blocks 0 .. tb_per + Q/KT + 2 of a sub-plane), not the whole plane -- with
    // the compile-time stride the planes are larger than one window
    const int need = KT * (a.tb_per + Q / KT + 3);
    const int nfr = (need < plane_len ? need : plane_len) * 4 / CH;
    const int c_base = K * KT * tb_lo;               // first candidate of this CTA
    int ncand = K * KT * ntb; if (c_base + ncand > L) ncand = L - c_base; if (ncand < 0) ncand = 0;

    // ---- staging (all asynchronous) ----
    const auto window_copy = [&](long long pos_seq, float* Xb) {
        const long long p0 = pos_seq - a.in.prefix;
        for (int fr = tid; fr < nfr; fr += kTdsThreads) {
            const int f = f_lo + fr;
            const long long g = p0 + f;
            const bool ok = f < region && g >= 0 && g < a.in.n;
            const long long go = ok ? g : 0;
            if (CH == 2) {
                const int m = fr >> 1, pl = (fr & 1) * 2;
                float* d = Xb + pl * plane_len + (m % KT) * sk + m / KT;
                cp_async4(d, base.b ? base.a + go : base.a + 2 * go, ok);
                cp_async4(d + plane_len, base.b ? base.b + go : base.a + 2 * go + 1, ok);
            } else {
                const float* src = base.a + go;
                const int m = fr >> 2;
                cp_async4(Xb + (fr & 3) * plane_len + (m % KT) * sk + m / KT, src, ok);
            }
        }
    };
    const auto mid_region_copy = [&](long long first_frame) {
        const long long p0 = first_frame - a.in.prefix;
        for (int fr = tid; fr < mreg; fr += kTdsThreads) {
            const long long g = p0 + fr;
            const bool ok = g >= 0 && g < a.in.n;
            const long long go = ok ? g : 0;
            if (CH == 2) {
                if (base.b) { cp_async4(MR + 2 * fr, base.a + go, ok); cp_async4(MR + 2 * fr + 1, base.b + go, ok); }
                else if (a.vec8) cp_async8(MR + 2 * fr, base.a + 2 * go, ok);
                else { cp_async4(MR + 2 * fr, base.a + 2 * go, ok); cp_async4(MR + 2 * fr + 1, base.a + 2 * go + 1, ok); }
            } else cp_async4(MR + fr, base.a + go, ok);
        }
    };
    const auto l2_prefetch = [&](long long pos_seq) {
        const long long q0 = pos_seq + f_lo - a.in.prefix;
        const int lines = (nfr * CH * 4 + 127) / 128 + 1;
        for (int t = tid; t < lines; t += blockDim.x) {
            const long long f = q0 + (long long)t * (32 / CH);
            if (f >= 0 && f < a.in.n) {
                if (CH == 2 && base.b) {
                    // planes: a 128-byte line holds 32 frames of one plane; t walks 16-frame steps, so alternate the planes
                    asm volatile("prefetch.global.L2 [%0];" :: "l"((t & 1 ? base.b : base.a) + f));
                } else asm volatile("prefetch.global.L2 [%0];" :: "l"(base.a + f * CH));
            }
        }
    };
    // norm units: thread = (plane rho, KT consecutive start positions): N[rho][m] = sum_q X_rho[m + q]^2, which is
    // the norm lane sum of every (l, kappa, t) with plane (CH*kappa + l) & 3 and t + ((CH*kappa + l) >> 2) == m --
    // same samples, same order, computed once.
    const auto norm_units = [&](const float* X) {
        for (int v = warp; v < nnorm; v += nwarps) {
            const int rho = v / wpn, mb = (v - rho * wpn) * 32 + lane;
            if (mb <= ntb) {
                float nr[KT];
                bool summed = false;
                if constexpr (kAhead) { if (Q % KT == 0) { tds_lane_sums_ahead<KT, 1, SK>(X + rho * plane_len + mb, sk, nullptr, Q, nr); summed = true; } }
                if (!summed) tds_lane_sums<KT, 1, SK, ROT>(X + rho * plane_len + mb, sk, nullptr, Q, nr);
#pragma unroll
                for (int k = 0; k < KT; k++) PN[rho * a.npm + KT * mb + k + mb * ((KT & 1) ? 2 : 1)] = nr[k];      // = ps_slot<KT>(KT * mb + k)
            }
        }
    };

    if (a.seq_end <= a.seq_begin) return;            // uniform over the cluster: nobody touches a peer
    if (CL > 1) {
        if (tid == 0) {
            st_mbar_init(&xbar[0], 1); st_mbar_init(&xbar[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster.sync();                              // every CTA's barriers exist before a peer signals them
    }

    // position weights (sequence independent): 1 - 0.25 t^2, t = (2c - L) / L
    for (int cc = tid; cc < ncand; cc += blockDim.x) {
        const int c = c_base + cc;
        const double tmp = __ddiv_rn((double)(2 * c - L), (double)L);
        PW[cc] = __dsub_rn(1.0, __dmul_rn(__dmul_rn(0.25, tmp), tmp));
        // where the candidate's correlation lane sum and its four norm lane sums live: sequence independent, so the
        // divisions by K*KT and KT of the padded layouts are done once per launch, not once per candidate and sequence
        SLT[5 * cc] = ps_slot<K * KT>(cc);
        const int kap = cc % K, trel = cc / K;
#pragma unroll
        for (int l = 0; l < 4; l++) {
            const int u0 = CH * kap + l;
            SLT[5 * cc + 1 + l] = (u0 & 3) * a.npm + ps_slot<KT>(trel + (u0 >> 2));
        }
    }
    // sequence positions ride two iterations ahead in registers: a load of pos[] at the top of an iteration put an L2 round
    // trip on the chain's critical path (phase timers: 0.5 k of the 12.7 k clocks of a sequence at 32 tracks)
    long long pos_cur = a.pos[a.seq_begin];
    long long pos_n1 = a.seq_begin + 1 < a.seq_end ? a.pos[a.seq_begin + 1] : 0;
    long long pos_n2 = a.seq_begin + 2 < a.seq_end ? a.pos[a.seq_begin + 2] : 0;
    window_copy(pos_cur, X0);
    int moff = 0;                                    // offset of the mid buffer inside the staged region
    if (a.seq_begin == 1) mid_region_copy(a.pos[0] + temp);       // first sequence: offset 0, no search
    else {
        // continuing a chain: the mid buffer was cut at the previous sequence's offset (written by the launch before)
        mid_region_copy(a.pos[a.seq_begin - 1] + ovl + temp);
        moff = offs[a.seq_begin - 2];
    }
    cp_async_wait_all();
    __syncthreads();
    norm_units(X0);
    int cur = 0;
#ifdef NODEY_TDS_TIMING
    const bool timing_on = blockIdx.x == 0 && tid == 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
#endif

    for (int i = a.seq_begin; i < a.seq_end; i++) {
        const long long p0 = pos_cur;
        const float* X = X0 + cur * 4 * plane_len;
        // ---- mid buffer: gathered from the staged region, de-interleaved by lane ----
        for (int j = tid; j < 4 * Q; j += blockDim.x) {
            const int q = j >> 2;
            Y[(j & 3) * QP + (KTP == KT ? q : (q / KT) * KTP + q % KT)] = MR[moff * CH + j];
        }
        if (i + 1 < a.seq_end) window_copy(pos_n1, X0 + (cur ^ 1) * 4 * plane_len);
        if (i + 2 < a.seq_end) l2_prefetch(pos_n2);
        pos_cur = pos_n1; pos_n1 = pos_n2;
        if (i + 3 < a.seq_end) pos_n2 = a.pos[i + 3];
        TDS_T(0);
        __syncthreads();
        TDS_T(1);
        if (i + 1 < a.seq_end) mid_region_copy(p0 + ovl + temp);  // the region is free again: everybody has gathered

        // ---- correlation lane sums: thread = (lane l, class kappa, KT consecutive candidates of the class) ----
        for (int unit = warp; unit < nunits; unit += nwarps) {
            const int combo = unit / wpc, wsub = unit - combo * wpc;
            const int l = combo & 3, kappa = combo >> 2;
            const int tb = wsub * 32 + lane;
            // stream of (l, kappa): plane (CH*kappa + l) & 3, starting `off` = 0 or 1 elements behind the candidate's index.
            // Slot k of thread tb walks the elements KT*tb + k + s; it belongs to candidate KT*tb + k - off of the class,
            // so that both kinds of stream run the same code with the same immediate offsets.
            const int u0 = CH * kappa + l, off = u0 >> 2;
            if (tb < ntb + off) {
                const float* xb = X + (u0 & 3) * plane_len + tb;
                const float* yp = Y + l * QP;
                float acc[KT];
                bool summed = false;
                if constexpr (kAhead) { if (Q % KT == 0) { tds_lane_sums_ahead<KT, 0, SK>(xb, sk, yp, Q, acc); summed = true; } }
                if (!summed) tds_lane_sums<KT, 0, SK, ROT>(xb, sk, yp, Q, acc);
                // candidate cc = K*KT*tb + (kappa + K*(k - off)): its pad count cc / (K*KT) is tb, or tb - 1 for the one slot of a
                // late stream that belongs to the block before (k = 0, off = 1) -- no division per store
                constexpr int S = K * KT, PADW = (S & 1) ? 2 : 1;
                const int cc0 = S * tb + kappa - K * off;
                float* ps_row = PS + l * a.ncand_pad + cc0 + PADW * tb;
#pragma unroll
                for (int k = 0; k < KT; k++) {
                    const int tl = KT * tb + k - off;
                    const int cc = cc0 + K * k;
                    if (tl >= 0 && cc < ncand) ps_row[K * k - (k == 0 ? PADW * off : 0)] = acc[k];
                }
            }
        }
        TDS_T(2);
        __syncthreads();
        TDS_T(3);

        // ---- per candidate: horizontal add in the SSE order, normalise, weight; arg-max (first wins) ----
        unsigned long long bk = 0ull; int bi = 0x7fffffff;
        for (int cc = tid; cc < ncand; cc += blockDim.x) {
            const int np = a.ncand_pad, sl = SLT[5 * cc];
            const float sum = __fadd_rn(__fadd_rn(__fadd_rn(PS[sl], PS[np + sl]), PS[2 * np + sl]), PS[3 * np + sl]);
            float nl[4];
#pragma unroll
            for (int l = 0; l < 4; l++) nl[l] = PN[SLT[5 * cc + 1 + l]];
            const float nr = __fadd_rn(__fadd_rn(__fadd_rn(nl[0], nl[1]), nl[2]), nl[3]);
            const double dn = (double)nr;
            double corr = __ddiv_rn((double)sum, __dsqrt_rn(dn < 1e-9 ? 1.0 : dn));
            corr = __dmul_rn(__dadd_rn(corr, 0.1), PW[cc]);
            const unsigned long long k = argmax_key(corr);
            if (k > bk || bi == 0x7fffffff) { bk = k; bi = c_base + cc; }
        }
        {
            // warp arg-max with three integer reductions: high word, low word among the leaders, lowest index
            const unsigned hi = (unsigned)(bk >> 32), lo = (unsigned)bk;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            const int mi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? bi : 0x7fffffff);
            if (lane == 0) { red_k[warp] = ((unsigned long long)mh << 32) | ml; red_i[warp] = mi; }
        }
        TDS_T(4);
        __syncthreads();
        TDS_T(5);
        // every warp reduces the per-warp results itself, with the same three integer reductions (no serial walk over the
        // warps, no second barrier): largest key, lowest index among equals
        unsigned long long fk; int fi;
        {
            const unsigned long long k = lane < nwarps ? red_k[lane] : 0ull;
            const int ix = lane < nwarps ? red_i[lane] : 0x7fffffff;
            const unsigned hi = (unsigned)(k >> 32), lo = (unsigned)k;
            const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
            const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
            fi = __reduce_min_sync(0xffffffffu, (hi == mh && lo == ml) ? ix : 0x7fffffff);
            fk = ((unsigned long long)mh << 32) | ml;
        }
        // The CTAs of a cluster exchange their results WITHOUT a cluster barrier: each sends its (key, index) pair into its slot
        // in every CTA (its own included) with st.async, which counts the bytes on the RECEIVER's mbarrier; a CTA waits
        // until its CL pairs have landed.  (The barrier.cluster arrive with release semantics this replaces cost about 1 k
        // of a sequence's 12 k clocks at 32 tracks.)  Slots and barriers alternate by parity: a peer can only write parity
        // p again after it has seen this CTA's result of the sequence in between, which this CTA sends after reading p.
        const int it = i - a.seq_begin, par = it & 1;
        if (CL > 1) {
            if (tid == 0) st_mbar_arrive_expect(&xbar[par], CL * 16u);
            if (tid < (int)CL)
                st_async_pair(st_map_rank(&xch[par][crank][0], (unsigned)tid), fk, (unsigned long long)(unsigned)fi, st_map_rank(&xbar[par], (unsigned)tid));
        }
        TDS_T(6);
        // ---- in the shadow of the exchange: next window landed -> its norm sums ----
        cp_async_wait_all();
        __syncthreads();
        cur ^= 1;
        if (i + 1 < a.seq_end) norm_units(X0 + cur * 4 * plane_len);
        if (CL > 1) {
            st_mbar_wait_cluster(&xbar[par], (unsigned)(it >> 1) & 1u);
            fk = xch[par][0][0]; fi = (int)(unsigned)xch[par][0][1];
            for (unsigned r = 1; r < CL; r++) {
                const unsigned long long k = xch[par][r][0]; const int ix = (int)(unsigned)xch[par][r][1];
                if (k > fk || (k == fk && ix < fi)) { fk = k; fi = ix; }
            }
        }
        TDS_T(7);
        if (fi == 0x7fffffff) fi = 0;
        if (crank == 0 && tid == 0) offs[i - 1] = fi;
        moff = fi;
        // the slots of parity `par` are rewritten two sequences later (see above);
        // Y/PS/MR are rewritten only after the next __syncthreads of this CTA, PN after the one above
    }
    if (CL > 1) cluster.sync();      // no CTA may exit while a peer can still write into its shared memory
#ifdef NODEY_TDS_TIMING
    if (timing_on)
        for (int k = 0; k < 8; k++) g_tds_phase[k] = (unsigned long long)tacc[k];
#endif
}

// eight CTAs per track (KT = 2, four CTAs per SM) up to this many tracks per launch on a 148-SM GPU.  Measured in the engine
// (bench.py --tracks n, both nodes' chains side by side): 32 tracks 41.0 ms against 31.1 ms with four CTAs per track, 16
// tracks 27.0 against 23.0 ms -- a sequence's fixed costs (gather, exchange, barriers) do not shrink with the slice and two
// accumulators per thread leave the FP32 pipe waiting on its own latency
constexpr long long kTdsCluster8Tracks = 0;       // measured: never; the variant stays selectable for tests
constexpr int kTdsSkPair = 132;      // compile-time sub-plane stride of the KT = 2 kernel (eight CTAs per track: 29 candidate blocks + 96 groups + margin; 4 mod 8)
constexpr int kTdsSkLarge = 51;      // compile-time sub-plane stride of the KT = 11..16 kernels (odd; 32 candidate blocks + 13 groups + margin)
typedef void (*TdsKernel)(TdsArgs);
// KT = 11..16 exist for stereo with the compile-time stride only (mono streams fit one warp at KT = 8; anything that does
// not fit the fixed stride runs with KT = 8 and the run-time stride)
template <int CH>
static TdsKernel tds_kernel_ch(int KT, bool fixed)
{
    switch (KT) {
    case 2:  return fixed ? tds_offsets_kernel<CH, 2, kTdsSkPair> : tds_offsets_kernel<CH, 2, 0>;
    case 4:  return fixed ? tds_offsets_kernel<CH, 4, 84> : tds_offsets_kernel<CH, 4, 0>;
    case 8:  return fixed ? tds_offsets_kernel<CH, 8, 92> : tds_offsets_kernel<CH, 8, 0>;
    }
    if constexpr (CH == 2) {
        if (!fixed) return nullptr;
        switch (KT) {
        case 11: return tds_offsets_kernel<CH, 11, kTdsSkLarge>;
        case 12: return tds_offsets_kernel<CH, 12, kTdsSkLarge>;
        case 13: return tds_offsets_kernel<CH, 13, kTdsSkLarge>;
        case 14: return tds_offsets_kernel<CH, 14, kTdsSkLarge>;
        case 15: return tds_offsets_kernel<CH, 15, kTdsSkLarge>;
        case 16: return tds_offsets_kernel<CH, 16, kTdsSkLarge>;
        }
    }
    return nullptr;
}
static TdsKernel tds_kernel_for(int CH, int KT, bool fixed) { return CH == 2 ? tds_kernel_ch<2>(KT, fixed) : tds_kernel_ch<1>(KT, fixed); }

// ---------------------------------------------------------------------------------------------
// TDStretch assemble: overlap (cross-fade) + copy of every sequence, one CTA per (sequence, track)
// ---------------------------------------------------------------------------------------------
struct AsmArgs {
    View in; TrackTab tt;
    float* out; long long out_stride; long long out_cap;   // frames to write at most (trim)
    const long long* pos;
    const int* offs; long long offs_stride;
    const float* fade;          // stereo: [2][overlap] (f1, f2)
    int nseq, overlap, seek_window;
};

template <int CH>
__global__ void __launch_bounds__(256) tds_assemble_kernel(const __grid_constant__ AsmArgs a)
{
    const int i = blockIdx.x;
    const long long track = blockIdx.y;
    const Src base = view_src(a.in, a.tt, track);
    float* out = a.out + track * a.out_stride;
    const int* offs = a.offs + track * a.offs_stride;
    const int ovl = a.overlap, temp = a.seek_window - 2 * ovl;
    if (i == 0) {
        const long long p0 = a.pos[0];
        for (int k = threadIdx.x; k < temp; k += blockDim.x) {
            if (k >= a.out_cap) break;
#pragma unroll
            for (int c = 0; c < CH; c++) out[(long long)k * CH + c] = view_sample<CH>(a.in, base, p0 + k, c);
        }
        return;
    }
    const long long ob = temp + (long long)(i - 1) * (a.seek_window - ovl);
    const long long src = a.pos[i] + offs[i - 1];
    const long long mid = (i == 1) ? a.pos[0] + temp : a.pos[i - 1] + offs[i - 2] + ovl + temp;
    for (int k = threadIdx.x; k < ovl + temp; k += blockDim.x) {
        const long long o = ob + k;
        if (o >= a.out_cap) break;
        if (k < ovl) {
            if (CH == 2) {
                const float f1 = a.fade[k], f2 = a.fade[ovl + k];
#pragma unroll
                for (int c = 0; c < 2; c++) {
                    const float x = view_sample<2>(a.in, base, src + k, c), m = view_sample<2>(a.in, base, mid + k, c);
                    out[o * 2 + c] = __fadd_rn(__fmul_rn(x, f1), __fmul_rn(m, f2));
                }
            } else {
                const float x = view_sample<1>(a.in, base, src + k, 0), m = view_sample<1>(a.in, base, mid + k, 0);
                out[o] = __fdiv_rn(__fadd_rn(__fmul_rn(x, (float)k), __fmul_rn(m, (float)(ovl - k))), (float)ovl);
            }
        } else {
#pragma unroll
            for (int c = 0; c < CH; c++) out[o * CH + c] = view_sample<CH>(a.in, base, src + k, c);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// AAFilter FIR (64 taps).  Tile of outputs per CTA, input tile in shared memory, R outputs/thread.
// ---------------------------------------------------------------------------------------------
struct FirArgs {
    View in;
    float* out; long long out_stride;
    long long count;            // outputs per track
    float h[kAaLen];            // coefficients ride in the kernel parameters (constant bank, uniform reads)
};

constexpr int kFirR = 4;                  // outputs per thread (mono path)
constexpr int kFirThreads = 256;
constexpr int kFirTile = kFirR * kFirThreads;   // 1024 outputs per CTA (mono path)

// mono: FIRFilter::evaluateFilterMono, double accumulator
__global__ void __launch_bounds__(kFirThreads) aa_fir_mono_kernel(const __grid_constant__ FirArgs a)
{
    __shared__ __align__(16) float tile[kFirTile + kAaLen];
    const long long track = blockIdx.y;
    const Src base = view_src(a.in, track);
    float* out = a.out + track * a.out_stride;
    const float* h = a.h;
    const long long ntiles = (a.count + kFirTile - 1) / kFirTile;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long n0 = t * kFirTile;
        __syncthreads();
        for (int f = threadIdx.x; f < kFirTile + kAaLen; f += blockDim.x) tile[f] = view_sample<1>(a.in, base, n0 + f, 0);
        __syncthreads();
        double s[kFirR];
#pragma unroll
        for (int r = 0; r < kFirR; r++) s[r] = 0.0;
        const float* tp = tile + threadIdx.x;
        for (int i = 0; i < kAaLen; i++) {
            const float hi = h[i];
#pragma unroll
            for (int r = 0; r < kFirR; r++) s[r] = __dadd_rn(s[r], (double)__fmul_rn(tp[r * kFirThreads + i], hi));
        }
#pragma unroll
        for (int r = 0; r < kFirR; r++) {
            const long long n = n0 + threadIdx.x + r * kFirThreads;
            if (n < a.count) out[n] = (float)s[r];
        }
    }
}

// stereo: FIRFilterSSE::evaluateFilterStereo.  A thread owns 8 CONSECUTIVE output frames and slides one
// register window of 10 input frames over the 32 tap pairs: one 128-bit shared load (two new frames)
// feeds 64 rounded multiply/adds.  The tile is stored in 16-byte chunks (two frames) whose column is
// XOR-swizzled with the 128-byte row, so the threads' strided 128-bit loads are conflict free.
constexpr int kFirS = 8;                              // outputs per thread
constexpr int kFirTileS = kFirS * kFirThreads;        // 2048 outputs per CTA
constexpr int kFirChunks = (kFirTileS + kAaLen + 8) / 2;
static_assert(kFirThreads % 128 == 0, "st_post advances tile slots by kFirThreads frames: the chunk swizzle must repeat");

__device__ __forceinline__ int fir_chunk(int chunk) { return (chunk & ~7) | ((chunk ^ (chunk >> 3)) & 7); }

__global__ void __launch_bounds__(kFirThreads) aa_fir_stereo_kernel(const __grid_constant__ FirArgs a)
{
    __shared__ __align__(16) float4 tile[kFirChunks];
    const long long track = blockIdx.y;
    const Src base = view_src(a.in, track);
    float* out = a.out + track * a.out_stride;
    const float* h = a.h;
    const long long ntiles = (a.count + kFirTileS - 1) / kFirTileS;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long n0 = t * kFirTileS;
        __syncthreads();
        for (int c = threadIdx.x; c < kFirChunks; c += blockDim.x) {
            const float2 f0 = view_frame2<2>(a.in, base, n0 + 2 * c), f1 = view_frame2<2>(a.in, base, n0 + 2 * c + 1);
            tile[fir_chunk(c)] = make_float4(f0.x, f0.y, f1.x, f1.y);
        }
        __syncthreads();
        float e0[kFirS], e1[kFirS], o0[kFirS], o1[kFirS];
#pragma unroll
        for (int r = 0; r < kFirS; r++) { e0[r] = e1[r] = o0[r] = o1[r] = 0.f; }
        // window: frames 8*tid + 2*step .. + 9, kept as 5 chunks in a rotating register file
        float4 w[5];
        const int c0 = threadIdx.x * (kFirS / 2);
#pragma unroll
        for (int k = 0; k < 5; k++) w[k] = tile[fir_chunk(c0 + k)];
#pragma unroll
        for (int step = 0; step < kAaLen / 2; step++) {
            const float h0 = h[2 * step], h1 = h[2 * step + 1];
#pragma unroll
            for (int r = 0; r < kFirS; r++) {
                // even tap: frame r of the window; odd tap: frame r + 1
                const float4 ce = w[(step + r / 2) % 5], co = w[(step + (r + 1) / 2) % 5];
                const float xe0 = (r & 1) ? ce.z : ce.x, xe1 = (r & 1) ? ce.w : ce.y;
                const float xo0 = ((r + 1) & 1) ? co.z : co.x, xo1 = ((r + 1) & 1) ? co.w : co.y;
                e0[r] = __fadd_rn(e0[r], __fmul_rn(xe0, h0));
                e1[r] = __fadd_rn(e1[r], __fmul_rn(xe1, h0));
                o0[r] = __fadd_rn(o0[r], __fmul_rn(xo0, h1));
                o1[r] = __fadd_rn(o1[r], __fmul_rn(xo1, h1));
            }
            if (step + 1 < kAaLen / 2) w[step % 5] = tile[fir_chunk(c0 + step + 5)];    // the chunk that just left the window
        }
        const long long n = n0 + (long long)threadIdx.x * kFirS;
        float4* o4 = reinterpret_cast<float4*>(out + 2 * n);       // track strides and n are multiples of 2 frames: 16-byte aligned
#pragma unroll
        for (int r = 0; r < kFirS; r += 2) {
            const float4 v = make_float4(__fadd_rn(o0[r], e0[r]), __fadd_rn(o1[r], e1[r]), __fadd_rn(o0[r + 1], e0[r + 1]), __fadd_rn(o1[r + 1], e1[r + 1]));
            if (n + r + 1 < a.count) o4[r / 2] = v;
            else if (n + r < a.count) { out[2 * (n + r)] = v.x; out[2 * (n + r) + 1] = v.y; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cubic transposer: output i reads 4 frames at floor(i*rate), fraction (float)frac(i*rate)
// ---------------------------------------------------------------------------------------------
struct CubicArgs {
    View in; TrackTab tt;
    float* out; long long out_stride;
    long long count;
    unsigned long long R;   // rate = R * 2^-e
    int e;
};

template <int CH>
__global__ void __launch_bounds__(256) cubic_kernel(const __grid_constant__ CubicArgs a)
{
    const long long track = blockIdx.y;
    const Src base = view_src(a.in, a.tt, track);
    float* out = a.out + track * a.out_stride;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double inv = 1.0 / (double)(1ull << a.e);     // exact power of two (e <= 62)
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += stride) {
        const unsigned long long lo = (unsigned long long)i * a.R, hi = __umul64hi((unsigned long long)i, a.R);
        const long long P = (long long)((lo >> a.e) | (a.e ? (hi << (64 - a.e)) : 0ull));
        const unsigned long long fb = lo & ((1ull << a.e) - 1ull);
        const float x2 = (float)((double)fb * inv);
        const float x1 = __fmul_rn(x2, x2);
        const float x0 = __fmul_rn(x1, x2);
        // y_r = ((c0*x0 + c1*x1) + c2*x2) + c3*1, coefficients of InterpolateCubic.cpp
        const float y0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-0.5f, x0), __fmul_rn(1.0f, x1)), __fmul_rn(-0.5f, x2)), __fmul_rn(0.0f, 1.0f));
        const float y1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(1.5f, x0), __fmul_rn(-2.5f, x1)), __fmul_rn(0.0f, x2)), __fmul_rn(1.0f, 1.0f));
        const float y2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-1.5f, x0), __fmul_rn(2.0f, x1)), __fmul_rn(0.5f, x2)), __fmul_rn(0.0f, 1.0f));
        const float y3 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(0.5f, x0), __fmul_rn(-0.5f, x1)), __fmul_rn(0.0f, x2)), __fmul_rn(0.0f, 1.0f));
        if (CH == 2) {
            const float2 p0 = view_frame2<2>(a.in, base, P), p1 = view_frame2<2>(a.in, base, P + 1);
            const float2 p2 = view_frame2<2>(a.in, base, P + 2), p3 = view_frame2<2>(a.in, base, P + 3);
            float2 o;
            o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(y0, p0.x), __fmul_rn(y1, p1.x)), __fmul_rn(y2, p2.x)), __fmul_rn(y3, p3.x));
            o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(y0, p0.y), __fmul_rn(y1, p1.y)), __fmul_rn(y2, p2.y)), __fmul_rn(y3, p3.y));
            reinterpret_cast<float2*>(out)[i] = o;
        } else {
            const float p0 = view_sample<1>(a.in, base, P, 0), p1 = view_sample<1>(a.in, base, P + 1, 0);
            const float p2 = view_sample<1>(a.in, base, P + 2, 0), p3 = view_sample<1>(a.in, base, P + 3, 0);
            out[i] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(y0, p0), __fmul_rn(y1, p1)), __fmul_rn(y2, p2)), __fmul_rn(y3, p3));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fused tail of the TDStretch-first order (rate > 1, stereo): cross-fade + sequence copy -> AA FIR ->
// cubic transposer in one kernel.  A CTA owns 2040 FIR outputs (computes 2048: the cubic kernel needs
// three frames beyond its last read position): it assembles the TDStretch output it needs straight
// from the node's input and the offset trace into the swizzled shared tile, runs the sliding-window
// FIR, parks the filtered frames in shared memory and interpolates every output whose read position
// falls into the tile.  Same rounding sequence as the three separate kernels; their two intermediate
// streams (2 x 8 B per frame, written and read) never touch HBM.
// ---------------------------------------------------------------------------------------------
// Packed multiply and add, each rounded separately, both channels of a frame in one instruction.  ptxas (12.9) contracts a
// mul.rn.f32x2 feeding an add.rn.f32x2 into ONE fma.rn.f32x2 -- explicit .rn and --fmad false notwithstanding, and also
// when the two are spelled fma(x, y, -0) and fma(a, 1, b) with literal constants (measured: the fused tail came out 1 ulp
// off in half of its samples; SASS showed one FFMA2 per tap).  The scalar forms are protected against that.  So the two
// neutral operands are RUN-TIME values (kernel parameters -0.0f and 1.0f): the assembler cannot know them, cannot simplify
// the two fused multiply-adds and cannot merge them, and the bits are those of the separate operations:
//   x * y = fma(x, y, -0)   one rounding of the exact product; adding -0 keeps the sign of a zero product
//   a + b = fma(a, 1, b)    a * 1 is exact
// tools/micro/f32x2_check.cu compares both with the scalar operations on the GPU (signed zeros, denormals, infinities).
struct Neutral { float neg_zero, one; };
__device__ __forceinline__ float2 mul2_rn(float2 x, float2 y, const Neutral& u) { return __ffma2_rn(x, y, make_float2(u.neg_zero, u.neg_zero)); }
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b, const Neutral& u) { return __ffma2_rn(a, make_float2(u.one, u.one), b); }

struct PostArgs {
    View in; TrackTab tt;
    const long long* pos; const int* offs; long long offs_stride;
    const float* fade;
    int nseq, overlap, seek_window, prefill;
    long long l1;               // TDStretch output frames
    float h[kAaLen];
    unsigned long long R; int e;
    long long tile_outputs;                                 // floor(kPostStride * 2^e / R): cubic outputs per tile, or one more
    float* out; long long out_stride; long long count;     // final frames to write
    long long tile_begin, tile_end;                         // this launch covers tiles [tile_begin, tile_end)
    Neutral u;                                              // -0.0f and 1.0f as run-time values (see mul2_rn / add2_rn)
};

constexpr int kPostStride = kFirTileS - 8;        // FIR outputs a tile contributes to the cubic stage

// read position of cubic output i: floor(i * R / 2^e)
__device__ __forceinline__ long long cubic_pos(unsigned long long i, unsigned long long R, int e)
{
    const unsigned long long lo = i * R, hi = __umul64hi(i, R);
    return (long long)((lo >> e) | (e ? (hi << (64 - e)) : 0ull));
}

// smallest i with cubic_pos(i) >= n, i.e. ceil(n * 2^e / R): double estimate, exact correction (a 128-bit
// division would be a several-hundred-instruction software loop)
__device__ __forceinline__ long long cubic_first_at(long long n, unsigned long long R, int e, double inv_rate)
{
    if (n <= 0) return 0;
    // the estimate is floor-ish: one above it is the answer nearly always, which the two loops then confirm with one
    // evaluation each (starting AT the estimate took three)
    long long i = (long long)((double)n * inv_rate) + 1;
    if (i < 0) i = 0;
    while (cubic_pos((unsigned long long)i, R, e) < n) i++;
    while (i > 0 && cubic_pos((unsigned long long)(i - 1), R, e) >= n) i--;
    return i;
}

__global__ void __launch_bounds__(kFirThreads) st_post_kernel(const __grid_constant__ PostArgs a)
{
    __shared__ __align__(16) float4 tile[kFirChunks];
    __shared__ float2 filt[kFirTileS + kFirTileS / 8];     // one pad slot per 8 frames: thread-strided stores stay 2-way
    const long long track = blockIdx.y;
    const Src base = view_src(a.in, a.tt, track);
    const int* offs = a.offs + track * a.offs_stride;
    float* out = a.out + track * a.out_stride;
    const int ovl = a.overlap, temp = a.seek_window - 2 * ovl, hop = a.seek_window - ovl;

    // TDStretch output frame o (0 <= o < l1): position k of a sequence whose input source (src) and
    // cross-fade partner (mid) positions are given; seq0 = first sequence (plain copy from src)
    const auto tds_frame = [&](long long o, bool seq0, long long src, long long mid, int k) -> float2 {
        if (o < 0 || o >= a.l1) return make_float2(0.f, 0.f);
        const float2 x = view_frame2<2>(a.in, base, src + k);
        if (seq0 || k >= ovl) return x;
        const float2 m = view_frame2<2>(a.in, base, mid + k);
        const float f1 = a.fade[k], f2 = a.fade[ovl + k];
        return make_float2(__fadd_rn(__fmul_rn(x.x, f1), __fmul_rn(m.x, f2)), __fadd_rn(__fmul_rn(x.y, f1), __fmul_rn(m.y, f2)));
    };
    // source / partner positions of sequence i (i >= 1); sequence 0 copies from pos[0]
    const auto seq_src = [&](int i) { return i < a.nseq ? a.pos[i] + offs[i - 1] : 0ll; };
    const auto seq_mid = [&](int i) { return i >= a.nseq ? 0ll : (i == 1 ? a.pos[0] + temp : a.pos[i - 1] + offs[i - 2] + ovl + temp); };

    const double inv_rate = (double)(1ull << a.e) / (double)a.R;
    // tiles until the last cubic read position is covered
    for (long long t = a.tile_begin + blockIdx.x; t < a.tile_end; t += gridDim.x) {
        const long long n0 = t * kPostStride;
        // cubic outputs whose read position lies in [n0, n0 + kPostStride): i in [i_lo, i_hi)
        // i_lo = ceil(n0 * 2^e / R) by an exact search; i_hi = ceil((n0 + stride) * 2^e / R) is i_lo + floor(stride * 2^e / R)
        // or one more (difference of two ceilings): one evaluation decides
        const long long i_lo = cubic_first_at(n0, a.R, a.e, inv_rate);
        long long i_hi = i_lo + a.tile_outputs;
        if (cubic_pos((unsigned long long)i_hi, a.R, a.e) < n0 + kPostStride) i_hi++;
        if (i_lo >= a.count) break;
        if (i_hi > a.count) i_hi = a.count;
        __syncthreads();
        {
            // sequence / position of the tile's first frame: one division per tile; the tile spans at most
            // kMaxSeq sequences whose source positions are fetched once (generic stepping beyond that)
            constexpr int kMaxSeq = 3;
            const long long o_first = n0 - a.prefill;
            int i_first = 0; long long k_first = o_first;                   // inside sequence 0 (or before the stream)
            if (o_first >= temp) {
                // (a 64-bit division is a long dependent subroutine: 1.7 % of the kernel's instructions but 9 % of its stall
                // samples when every thread ran it per tile; streams shorter than 2^32 frames take the 32-bit one)
                const long long o2 = o_first - temp;
                const long long sq = o2 <= 0xffffffffll ? (long long)((unsigned)o2 / (unsigned)hop) : o2 / hop;
                i_first = 1 + (int)sq; k_first = o2 - sq * hop;
            }
            long long src_j[kMaxSeq], mid_j[kMaxSeq];
            const int i_base = i_first > 0 ? i_first : 1;
#pragma unroll
            for (int j = 0; j < kMaxSeq; j++) src_j[j] = seq_src(i_base + j);
            // the cross-fade partner of sequence i >= 2 starts ovl + temp behind the source of sequence i - 1 (seq_mid):
            // one pair of loads per sequence instead of two
            mid_j[0] = seq_mid(i_base);
#pragma unroll
            for (int j = 1; j < kMaxSeq; j++) mid_j[j] = i_base + j >= a.nseq ? 0ll : src_j[j - 1] + ovl + temp;
            const long long pos0 = a.pos[0];
            // interior tile: every frame it can touch lies inside the real input and inside sequences >= 1 that
            // were fetched above -> no bounds checks, 32-bit stepping
            const long long o_last = o_first + 2 * kFirChunks - 1;
            const long long reach = (long long)hop + 2 * kFirChunks + ovl;
            bool interior = i_first >= 1 && o_last < a.l1 && i_first + kMaxSeq - 1 < a.nseq && 2 * kFirChunks + k_first < (long long)kMaxSeq * hop;
#pragma unroll
            for (int j = 0; j < kMaxSeq; j++)
                interior = interior && src_j[j] - a.in.prefix >= 0 && src_j[j] - a.in.prefix + reach <= a.in.n;
            // partners: mid_j[j >= 1] = src_j[j - 1] + ovl + temp lies inside what the source checks cover (2 ovl + temp =
            // hop + ovl <= reach); the first one is on its own
            interior = interior && mid_j[0] - a.in.prefix >= 0 && mid_j[0] - a.in.prefix + ovl <= a.in.n;
            if (interior) {
                // The tile's frames fall into at most three sequences, each a cross-faded run of `ovl` frames followed by a
                // plain copy of `hop - ovl` frames: six runs with warp-uniform source pointers.  The copies (87 % of the
                // frames) go global -> shared with cp.async, no registers and no per-frame sequence logic; the per-frame
                // version of this loop was a quarter of the kernel's instructions (ncu, round 2).
                const float2* b2 = reinterpret_cast<const float2*>(base.a) - a.in.prefix;
                const float* pl = base.a - a.in.prefix;
                const float* pr = base.b - a.in.prefix;
                const bool planar = base.b != nullptr;
                const int kf = (int)k_first;
                constexpr int F = 2 * kFirChunks;
                float2* tile2 = reinterpret_cast<float2*>(tile);
                const auto slot = [](int f) { return 2 * fir_chunk(f >> 1) + (f & 1); };
#pragma unroll
                for (int j = 0; j < kMaxSeq; j++) {
                    const int fj = j * hop - kf;                                  // tile frame where sequence i_first + j starts
                    const long long so = src_j[j] - fj, mo = mid_j[j] - fj;       // source / partner frame of tile frame 0
                    const int cs = fj > 0 ? fj : 0, ce = fj + ovl < F ? fj + ovl : F;
                    // (a thread's frames are kFirThreads = 256 apart, and slot(f + 256) = slot(f) + 256: the swizzle looks at
                    // bits 1..6 of f only -- one slot computation per run, running pointers after that)
                    for (int f = cs + (int)threadIdx.x; f < ce; f += kFirThreads) {
                        const int k = f - fj;
                        const float2 x = planar ? make_float2(__ldg(pl + so + f), __ldg(pr + so + f)) : __ldg(b2 + so + f);
                        const float2 m = planar ? make_float2(__ldg(pl + mo + f), __ldg(pr + mo + f)) : __ldg(b2 + mo + f);
                        const float f1 = a.fade[k], f2 = a.fade[ovl + k];
                        tile2[slot(f)] = add2_rn(mul2_rn(x, make_float2(f1, f1), a.u), mul2_rn(m, make_float2(f2, f2), a.u), a.u);
                    }
                    const int ps = fj + ovl > 0 ? fj + ovl : 0, pe = fj + hop < F ? fj + hop : F;
                    const int f0 = ps + (int)threadIdx.x;
                    if (f0 < pe) {
                        float2* d = tile2 + slot(f0);
                        const int n_it = (pe - f0 + kFirThreads - 1) / kFirThreads;
                        if (planar) {
                            const float* sl = pl + so + f0;
                            const float* sr = pr + so + f0;
                            for (int it = 0; it < n_it; it++, d += kFirThreads, sl += kFirThreads, sr += kFirThreads) {
                                cp_async4(reinterpret_cast<float*>(d), sl, true);
                                cp_async4(reinterpret_cast<float*>(d) + 1, sr, true);
                            }
                        } else {
                            const float2* sp = b2 + so + f0;
                            for (int it = 0; it < n_it; it++, d += kFirThreads, sp += kFirThreads)
                                cp_async8(reinterpret_cast<float*>(d), reinterpret_cast<const float*>(sp), true);
                        }
                    }
                }
                cp_async_wait_all();
            } else
            for (int c = threadIdx.x; c < kFirChunks; c += blockDim.x) {
                float2 f[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const long long o = o_first + 2 * c + u;
                    int i = i_first; long long k = k_first + 2 * c + u;
                    if (i == 0 && o >= temp) { i = 1; k = o - temp; }
                    while (i > 0 && k >= hop) { k -= hop; i++; }
                    if (i == 0) f[u] = tds_frame(o, true, pos0, 0, (int)(o < 0 ? 0 : o));
                    else {
                        const int j = i - i_base;
                        long long src, mid;
                        if (j == 0) { src = src_j[0]; mid = mid_j[0]; }
                        else if (j == 1) { src = src_j[1]; mid = mid_j[1]; }
                        else if (j == 2) { src = src_j[2]; mid = mid_j[2]; }
                        else { src = seq_src(i); mid = seq_mid(i); }
                        f[u] = tds_frame(o, false, src, mid, (int)k);
                    }
                }
                tile[fir_chunk(c)] = make_float4(f[0].x, f[0].y, f[1].x, f[1].y);
            }
        }
        __syncthreads();
        {
            // Both channels of a frame meet the same tap, so the two rounded multiplies and the two rounded adds of a tap
            // are ONE packed instruction each (mul.rn.f32x2 / add.rn.f32x2: per component the same result as the scalar
            // operations, sm_100).  The FP32 pipe does the same work, but the kernel was ISSUE bound with 40 % of its issue
            // slots going to everything that is not FIR arithmetic (ncu, round 1: issue 88 %, FMA pipe 55 %): the packed
            // form needs half the slots for the FIR.
            float2 e[kFirS], o[kFirS];
#pragma unroll
            for (int r = 0; r < kFirS; r++) { e[r] = make_float2(0.f, 0.f); o[r] = make_float2(0.f, 0.f); }
            float4 w[5];
            const int c0 = threadIdx.x * (kFirS / 2);
#pragma unroll
            for (int k = 0; k < 5; k++) w[k] = tile[fir_chunk(c0 + k)];
#pragma unroll
            for (int step = 0; step < kAaLen / 2; step++) {
                const float h0 = a.h[2 * step], h1 = a.h[2 * step + 1];
                const float2 hh0 = make_float2(h0, h0), hh1 = make_float2(h1, h1);
#pragma unroll
                for (int r = 0; r < kFirS; r++) {
                    const float4 ce = w[(step + r / 2) % 5], co = w[(step + (r + 1) / 2) % 5];
                    const float2 xe = (r & 1) ? make_float2(ce.z, ce.w) : make_float2(ce.x, ce.y);
                    const float2 xo = ((r + 1) & 1) ? make_float2(co.z, co.w) : make_float2(co.x, co.y);
                    e[r] = add2_rn(e[r], mul2_rn(xe, hh0, a.u), a.u);
                    o[r] = add2_rn(o[r], mul2_rn(xo, hh1, a.u), a.u);
                }
                if (step + 1 < kAaLen / 2) w[step % 5] = tile[fir_chunk(c0 + step + 5)];
            }
            // park the filtered frames (frame q of the tile lives at q + q/8)
#pragma unroll
            for (int r = 0; r < kFirS; r++) filt[threadIdx.x * (kFirS + 1) + r] = add2_rn(o[r], e[r], a.u);
        }
        __syncthreads();
        // cubic transposer over the tile (InterpolateCubic::transposeStereo).  The read position of output i is i * R in
        // 2^-e fixed point (128 bits): one product per thread and tile, then a 128-bit add of blockDim * R per output.
        // Fraction: (float)((double)fb * 2^-e) with fb < 2^e <= 2^53 is ONE rounding of fb * 2^-e to 24 bits, which is
        // what the unsigned 64-bit -> float conversion does; the power-of-two scale afterwards is exact.
        // Coefficients: y_r = ((c0*x0 + c1*x1) + c2*x2) + c3 as InterpolateCubic.cpp spells them; products with the
        // constants 1.0f and 0.0f and the additions of the resulting +0 are dropped where that cannot change a bit: x2 is in
        // [0, 1), so x1, x0 >= +0, 0.0f * x2 = +0, and a + (+0) differs from a only for a = -0, which none of the partial
        // sums below can be (a sum is -0 only when both addends are; every sum here has an addend that is +0 or positive).
        const float inv_f = __uint_as_float((unsigned)(127 - a.e) << 23);          // 2^-e, e <= 60: a normal float
        const unsigned long long fmask = (1ull << a.e) - 1ull;
        // position = (tile-relative integer part q, fraction fb < 2^e): the product once, then q += step_int + carry and
        // fb = (fb + step_frac) mod 2^e per output -- the same integers as the 128-bit running sum, half the instructions
        int q; unsigned long long fb;
        {
            const unsigned long long i0 = (unsigned long long)(i_lo + threadIdx.x);
            const unsigned long long p_lo = i0 * a.R, p_hi = __umul64hi(i0, a.R);
            q = (int)((long long)((p_lo >> a.e) | (a.e ? (p_hi << (64 - a.e)) : 0ull)) - n0);
            fb = p_lo & fmask;
        }
        const unsigned long long step_lo = (unsigned long long)blockDim.x * a.R, step_hi = __umul64hi((unsigned long long)blockDim.x, a.R);
        const int step_int = (int)((step_lo >> a.e) | (a.e ? (step_hi << (64 - a.e)) : 0ull));
        const unsigned long long step_frac = step_lo & fmask;
        const int n_out = (int)(i_hi - i_lo);                  // at most tile_outputs + 1
        float2* op = reinterpret_cast<float2*>(out) + i_lo + threadIdx.x;
        for (int k = (int)threadIdx.x; k < n_out; k += (int)blockDim.x, op += blockDim.x) {
            const int q_cur = q;
            const unsigned long long fb_cur = fb;
            {
                const unsigned long long sum = fb + step_frac;            // < 2^(e+1), e <= 60: no overflow
                q += step_int + (int)(sum >> a.e);
                fb = sum & fmask;
            }
            const float x2 = __fmul_rn(__ull2float_rn(fb_cur), inv_f);
            const float x1 = __fmul_rn(x2, x2);
            const float x0 = __fmul_rn(x1, x2);
            // (a product with -c is the negated product with c, bit for bit: three multiplications serve six terms)
            const float h0 = __fmul_rn(0.5f, x0), h2 = __fmul_rn(0.5f, x2), t0 = __fmul_rn(1.5f, x0);
            const float y0 = __fadd_rn(__fadd_rn(-h0, x1), -h2);
            const float y1 = __fadd_rn(__fadd_rn(t0, __fmul_rn(-2.5f, x1)), 1.0f);
            const float y2 = __fadd_rn(__fadd_rn(-t0, __fmul_rn(2.0f, x1)), h2);
            const float y3 = __fadd_rn(h0, __fmul_rn(-0.5f, x1));
            const auto at = [&](int f) { return filt[f + (f >> 3)]; };
            const float2 p0 = at(q_cur), p1 = at(q_cur + 1), p2 = at(q_cur + 2), p3 = at(q_cur + 3);
            const float2 o = add2_rn(add2_rn(add2_rn(mul2_rn(p0, make_float2(y0, y0), a.u), mul2_rn(p1, make_float2(y1, y1), a.u), a.u),
                                             mul2_rn(p2, make_float2(y2, y2), a.u), a.u), mul2_rn(p3, make_float2(y3, y3), a.u), a.u);
            *op = o;
        }
    }
}

}  // namespace nodey

using namespace nodey;

// ---------------------------------------------------------------------------------------------
// host plan
// ---------------------------------------------------------------------------------------------
struct nodey_soundtouch {
    int sample_rate = 0, ch = 0;
    double rate = 1, tempo = 1;
    int td_first = 0;
    int overlap = 0, seek_window = 0, seek_length = 0, sample_req = 0;
    double nominal_skip = 0;
    unsigned long long R = 0; int e = 0;       // rate = R * 2^-e exactly
    int prefill = 0;                           // silent frames in front of the RateTransposer input
    int force_cluster = 0;                     // test hook: 0 = automatic cluster size for the offsets kernel
    int force_kt = 0;                          // test hook: 0 = automatic number of candidates per thread of the offsets kernel
    int force_unfused = 0;                     // test hook: separate assemble / FIR / cubic kernels
    float aa[kAaLen];
    float* d_fade = nullptr;
    int device = 0;                            // the plan's tables live on the device that was current at create
    // Sequence start positions on the device, one IMMUTABLE table per TDStretch input length: a table is uploaded
    // once on the stream of the call that needs it first and never rewritten or freed while the plan lives, so calls
    // on other streams (the Runner's lanes share cached plans) only have to wait for its upload event.
    struct PosTable { long long* d = nullptr; std::vector<long long> host; cudaEvent_t ready = nullptr; };
    std::map<long long, PosTable*> tables;
    // host-side length bookkeeping of a render (flush iterations included) by (input frames, putSamples chunk): a chunked
    // render asks for it once per launch, and walking the sequence table a few dozen times costs more than the launch
    struct LenPlan { long long n_ext, l1, l2, l3, nseq, tds_in, total; std::vector<long long> pos; };
    std::map<std::pair<long long, int>, std::shared_ptr<const LenPlan>> len_plans;
    std::mutex mu;
};

namespace {

long long tds_build_pos(nodey_soundtouch* s, long long n_in, std::vector<long long>& pos)
{
    // TDStretch::processSamples() input bookkeeping (offsets do not influence it)
    pos.clear();
    long long consumed = 0;
    double skip_fract = 0;
    int beginning = 1;
    while (n_in - consumed >= s->sample_req) {
        pos.push_back(consumed);
        if (beginning) {
            beginning = 0;
            const int skip = (int)(s->tempo * s->overlap + 0.5 * s->seek_length + 0.5);
            skip_fract -= skip;
            if (skip_fract <= -s->nominal_skip) skip_fract = -s->nominal_skip;
        }
        skip_fract += s->nominal_skip;
        const int ovl_skip = (int)skip_fract;
        skip_fract -= ovl_skip;
        if (ovl_skip >= n_in - consumed) consumed = n_in; else consumed += ovl_skip;
    }
    return (long long)pos.size();
}

long long tds_out_frames(const nodey_soundtouch* s, long long nseq)
{
    if (nseq <= 0) return 0;
    return (long long)(s->seek_window - 2 * s->overlap) + (nseq - 1) * (long long)(s->seek_window - s->overlap);
}

long long fir_count(const nodey_soundtouch* s, long long n)
{
    if (n < kAaLen) return 0;
    if (s->ch == 2) { const long long c = (n - kAaLen) & ~1ll; return c < 2 ? 0 : c; }
    return n - kAaLen;
}

// number of outputs i >= 0 with floor(i * rate) < n - 4
long long cubic_count(const nodey_soundtouch* s, long long n)
{
    const long long end = n - 4;
    if (end <= 0) return 0;
    // smallest i with i*R >= end * 2^e  ->  ceil(end * 2^e / R)
    const unsigned __int128 num = (unsigned __int128)end << s->e;
    return (long long)((num + s->R - 1) / s->R);
}

struct StageLens { long long n_ext, l1, l2, l3, nseq, tds_in; };

void stage_lengths(nodey_soundtouch* s, long long n_ext, StageLens* L, std::vector<long long>* pos_out)
{
    std::vector<long long> tmp;
    std::vector<long long>& pos = pos_out ? *pos_out : tmp;
    L->n_ext = n_ext;
    if (s->td_first) {
        L->tds_in = n_ext;
        L->nseq = tds_build_pos(s, n_ext, pos);
        L->l1 = tds_out_frames(s, L->nseq);
        L->l2 = fir_count(s, s->prefill + L->l1);
        L->l3 = cubic_count(s, L->l2);
    } else {
        L->l1 = cubic_count(s, s->prefill + n_ext);
        L->l2 = fir_count(s, L->l1);
        L->tds_in = L->l2;
        L->nseq = tds_build_pos(s, L->l2, pos);
        L->l3 = tds_out_frames(s, L->nseq);
    }
}

// SoundTouch::flush(): silent 128-frame blocks until the expected amount exists (at most 200)
long long plan_total(nodey_soundtouch* s, long long in_frames, int frame_size, StageLens* L, std::vector<long long>* pos)
{
    double expected = 0;
    const double denom = s->rate * s->tempo;
    for (long long p = 0; p < in_frames; p += frame_size) {
        const long long n = (in_frames - p) < frame_size ? (in_frames - p) : frame_size;
        expected += (double)n / denom;
    }
    const long long want = (long long)(long)(expected + 0.5);
    stage_lengths(s, in_frames, L, pos);
    if (L->l3 >= want) return L->l3;         // nothing left to flush: everything already emitted
    for (int k = 1; k <= 200; k++) {
        stage_lengths(s, in_frames + 128ll * k, L, pos);
        if (L->l3 >= want) return want;
    }
    return L->l3;
}

// plan_total() with its result remembered in the plan object (caller holds s->mu)
std::shared_ptr<const nodey_soundtouch::LenPlan> cached_lengths(nodey_soundtouch* s, long long in_frames, int frame_size)
{
    const auto key = std::make_pair(in_frames, frame_size);
    const auto it = s->len_plans.find(key);
    if (it != s->len_plans.end()) return it->second;
    auto p = std::make_shared<nodey_soundtouch::LenPlan>();
    StageLens L;
    p->total = plan_total(s, in_frames, frame_size, &L, &p->pos);
    p->n_ext = L.n_ext; p->l1 = L.l1; p->l2 = L.l2; p->l3 = L.l3; p->nseq = L.nseq; p->tds_in = L.tds_in;
    if (s->len_plans.size() >= 16) s->len_plans.clear();
    s->len_plans[key] = p;
    return p;
}

}  // namespace

extern "C" {

int nodey_soundtouch_create(nodey_soundtouch** out, int sample_rate, int channels, float rate_arg, float pitch_arg)
{
    NODEY_REQUIRE(out, NODEY_E_INVALID, "nodey_soundtouch_create: null out pointer");
    NODEY_REQUIRE(channels == 1 || channels == 2, NODEY_E_INVALID, "Unsupported channel count: %d", channels);
    // audio-velocity.cpp:371: SoundTouch is only fed 8 kHz .. 48 kHz
    NODEY_REQUIRE(sample_rate >= 8000 && sample_rate <= 48000, NODEY_E_RANGE,
                  "Unsupported sample rate %d (SoundTouch node accepts 8000..48000)", sample_rate);
    NODEY_REQUIRE(rate_arg > 0.f && pitch_arg > 0.f, NODEY_E_RANGE, "nodey_soundtouch_create: rate and pitch must be positive");
    nodey_soundtouch* s = new nodey_soundtouch();
    s->sample_rate = sample_rate; s->ch = channels;
    if (cudaGetDevice(&s->device) != cudaSuccess) { cudaGetLastError(); s->device = 0; }
    const double vrate = (double)rate_arg, vpitch = (double)pitch_arg;
    s->tempo = 1.0 / vpitch;              // virtualTempo = 1
    s->rate = vpitch * vrate;
    s->td_first = !(s->rate <= 1.0f);
    // TDStretch parameters (setParameters / calcSeqParameters / setTempo), App. B2
    int ovl = (sample_rate * 8) / 1000;
    if (ovl < 16) ovl = 16;
    ovl -= ovl % 8;
    s->overlap = ovl;
    double seq = (90.0 - ((40.0 - 90.0) / (2.0 - 0.5)) * 0.5) + ((40.0 - 90.0) / (2.0 - 0.5)) * s->tempo;
    seq = seq < 40.0 ? 40.0 : (seq > 90.0 ? 90.0 : seq);
    double seek = (20.0 - ((15.0 - 20.0) / (2.0 - 0.5)) * 0.5) + ((15.0 - 20.0) / (2.0 - 0.5)) * s->tempo;
    seek = seek < 15.0 ? 15.0 : (seek > 20.0 ? 20.0 : seek);
    s->seek_window = (sample_rate * (int)(seq + 0.5)) / 1000;
    if (s->seek_window < 2 * ovl) s->seek_window = 2 * ovl;
    s->seek_length = (sample_rate * (int)(seek + 0.5)) / 1000;
    s->nominal_skip = s->tempo * (s->seek_window - ovl);
    {
        const int intskip = (int)(s->nominal_skip + 0.5);
        const int a = intskip + ovl;
        s->sample_req = (a > s->seek_window ? a : s->seek_window) + s->seek_length;
    }
    // rate as an exact dyadic rational: frexp mantissa bits
    {
        int ex = 0;
        const double m = frexp(s->rate, &ex);            // rate = m * 2^ex, 0.5 <= m < 1
        unsigned long long R = (unsigned long long)ldexp(m, 53);   // 53-bit integer mantissa, exact
        int e = 53 - ex;
        while (e > 0 && (R & 1ull) == 0) { R >>= 1; e--; }
        // exactness of the reference's `fract += rate` in double needs (1 + rate) * 2^e < 2^53
        if (e < 0 || e > 60 || ldexp(1.0 + s->rate, e) >= 9007199254740992.0) {
            delete s;
            set_error("nodey_soundtouch_create: rate %.17g is outside the exactly representable range", vpitch * vrate);
            return NODEY_E_RANGE;
        }
        s->R = R; s->e = e;
    }
    s->prefill = (1 + 32) * 2 / channels;     // RateTransposer latency pre-fill, stored as stereo frames
    // AAFilter::calculateCoeffs(), 64 taps
    {
        const double cutoff = s->rate > 1.0 ? 0.5 / s->rate : 0.5 * s->rate;
        double work[kAaLen], sum = 0;
        const double wc = 2.0 * M_PI * cutoff;
        const double temp_coeff = (2 * M_PI) / (double)kAaLen;
        for (int i = 0; i < kAaLen; i++) {
            const double cnt = (double)i - (double)(kAaLen / 2);
            double temp = cnt * wc;
            const double h = (temp != 0) ? sin(temp) / temp : 1.0;
            const double w = 0.54 + 0.46 * cos(temp_coeff * cnt);
            temp = w * h;
            work[i] = temp;
            sum += temp;
        }
        const double scale = 16384.0f / sum;
        for (int i = 0; i < kAaLen; i++) {
            double temp = work[i] * scale;
            temp += (temp >= 0) ? 0.5 : -0.5;
            s->aa[i] = (float)temp / 16384.0f;
        }
    }
    // overlapStereo() fade ramps: repeated float adds of 1/overlap
    {
        std::vector<float> fade((size_t)2 * ovl);
        const float scale = 1.0f / (float)ovl;
        float f1 = 0, f2 = 1.0f;
        for (int k = 0; k < ovl; k++) { fade[(size_t)k] = f1; fade[(size_t)(ovl + k)] = f2; f1 += scale; f2 -= scale; }
        cudaError_t err = cudaMalloc((void**)&s->d_fade, sizeof(float) * fade.size());
        if (err == cudaSuccess) err = cudaMemcpy(s->d_fade, fade.data(), sizeof(float) * fade.size(), cudaMemcpyHostToDevice);
        if (err != cudaSuccess) { if (s->d_fade) cudaFree(s->d_fade); delete s; return cuda_fail(err, "fade table upload", __FILE__, __LINE__); }
    }
    *out = s;
    return NODEY_OK;
}

/* test hook: 1 = run cross-fade, FIR and cubic stage as separate kernels even where the fused one applies */
int nodey_soundtouch_set_unfused(nodey_soundtouch* s, int unfused)
{
    NODEY_REQUIRE(s, NODEY_E_INVALID, "nodey_soundtouch_set_unfused: null handle");
    s->force_unfused = unfused ? 1 : 0;
    return NODEY_OK;
}

#ifdef NODEY_TDS_TIMING
/* development build only: phase clocks of the last tds_offsets launch (thread 0 of CTA 0) */
int nodey_debug_tds_phases(unsigned long long out[8])
{
    NODEY_CUDA_OK(cudaDeviceSynchronize());
    NODEY_CUDA_OK(cudaMemcpyFromSymbol(out, nodey::g_tds_phase, sizeof(unsigned long long) * 8));
    return NODEY_OK;
}
#endif

/* test hook: force the cluster size (1, 2, 4 or 8; 0 = automatic) of the offsets kernel */
int nodey_soundtouch_set_cluster(nodey_soundtouch* s, int cluster)
{
    NODEY_REQUIRE(s && (cluster == 0 || cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8), NODEY_E_INVALID,
                  "nodey_soundtouch_set_cluster: cluster must be 0, 1, 2, 4 or 8");
    s->force_cluster = cluster;
    return NODEY_OK;
}

/* test hook: force the number of candidates a thread of the offsets kernel owns (2, 4, 8, 11..16; 0 = automatic) */
int nodey_soundtouch_set_candidates_per_thread(nodey_soundtouch* s, int kt)
{
    NODEY_REQUIRE(s && (kt == 0 || kt == 2 || kt == 4 || kt == 8 || (kt >= 11 && kt <= 16)), NODEY_E_INVALID,
                  "nodey_soundtouch_set_candidates_per_thread: must be 0, 2, 4, 8 or 11..16");
    s->force_kt = kt;
    return NODEY_OK;
}

void nodey_soundtouch_destroy(nodey_soundtouch* s)
{
    if (!s) return;
    if (s->d_fade) cudaFree(s->d_fade);
    for (auto& [n, t] : s->tables) { if (t->ready) cudaEventDestroy(t->ready); if (t->d) cudaFree(t->d); delete t; }
    delete s;
}

/* info_i[8]: overlap, seek_window, seek_length, sample_req, tdstretch_first, prefill, channels, sample_rate
 * info_d[3]: rate, tempo, nominal_skip */
int nodey_soundtouch_info(const nodey_soundtouch* s, int info_i[8], double info_d[3])
{
    NODEY_REQUIRE(s, NODEY_E_INVALID, "nodey_soundtouch_info: null handle");
    if (info_i) {
        info_i[0] = s->overlap; info_i[1] = s->seek_window; info_i[2] = s->seek_length; info_i[3] = s->sample_req;
        info_i[4] = s->td_first; info_i[5] = s->prefill; info_i[6] = s->ch; info_i[7] = s->sample_rate;
    }
    if (info_d) { info_d[0] = s->rate; info_d[1] = s->tempo; info_d[2] = s->nominal_skip; }
    return NODEY_OK;
}

int64_t nodey_soundtouch_out_frames(nodey_soundtouch* s, int64_t in_frames, int frame_size, int64_t* n_sequences)
{
    if (!s || in_frames < 0 || frame_size <= 0) return NODEY_E_INVALID;
    std::lock_guard<std::mutex> lock(s->mu);
    const std::shared_ptr<const nodey_soundtouch::LenPlan> lp = cached_lengths(s, in_frames, frame_size);
    if (n_sequences) *n_sequences = lp->nseq;
    return lp->total;
}

}  // extern "C"

namespace {

// the device copy of the sequence positions for a TDStretch input of `tds_in` frames (see nodey_soundtouch::tables)
int position_table(nodey_soundtouch* s, long long tds_in, const std::vector<long long>& pos, cudaStream_t st, const long long** d_out)
{
    auto it = s->tables.find(tds_in);
    if (it == s->tables.end()) {
        if (s->tables.size() >= 64) {
            // a plan that has seen this many different lengths: nothing may still be reading the old tables after a
            // device-wide wait, so they can go (never happens in a render: a cached plan is keyed by its input length)
            NODEY_CUDA_OK(cudaDeviceSynchronize());
            for (auto& [n, t] : s->tables) { if (t->ready) cudaEventDestroy(t->ready); if (t->d) cudaFree(t->d); delete t; }
            s->tables.clear();
        }
        auto* t = new nodey_soundtouch::PosTable();
        t->host = pos;
        if (t->host.empty()) t->host.push_back(0);
        cudaError_t e = cudaMalloc((void**)&t->d, sizeof(long long) * t->host.size());
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ready, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMemcpyAsync(t->d, t->host.data(), sizeof(long long) * t->host.size(), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(t->ready, st);
        if (e != cudaSuccess) {
            if (t->ready) cudaEventDestroy(t->ready);
            if (t->d) cudaFree(t->d);
            delete t;
            return cuda_fail(e, "sequence position table upload", __FILE__, __LINE__);
        }
        it = s->tables.emplace(tds_in, t).first;
    }
    NODEY_CUDA_OK(cudaStreamWaitEvent(st, it->second->ready, 0));     // no-op on the uploading stream, orders every other one
    *d_out = it->second->d;
    return NODEY_OK;
}

// How a render is cut into `nchunks` launches (fused stereo TDStretch-first path): chunk c searches sequences
// [seq_begin, seq_end) and then runs the tail over the tiles whose TDStretch frames are final by then.
struct ChunkPlan { int seq_begin, seq_end; long long tile_begin, tile_end, in_need, out_ready; };

long long post_tiles_total(const StageLens& L) { return (L.l2 + kPostStride - 1) / kPostStride + 1; }

bool chunkable(const nodey_soundtouch* s) { return s->td_first && s->ch == 2 && !s->force_unfused; }

int effective_chunks(const nodey_soundtouch* s, const StageLens& L, int want)
{
    if (!chunkable(s) || want <= 1) return 1;
    const long long nsearch = L.nseq - 1;
    long long n = nsearch / 32;                 // at least 32 sequences per launch
    if (n > want) n = want;
    return n < 1 ? 1 : (int)n;
}

void chunk_plan(const nodey_soundtouch* s, const StageLens& L, const std::vector<long long>& pos, long long in_frames, long long out_frames,
                int c, int nchunks, ChunkPlan* cp)
{
    const long long nsearch = L.nseq > 1 ? L.nseq - 1 : 0;
    const auto seq_end_of = [&](int k) { return (int)(1 + (nsearch * (k + 1) + nchunks - 1) / nchunks); };
    const long long tiles = post_tiles_total(L);
    const int temp = s->seek_window - 2 * s->overlap, hop = s->seek_window - s->overlap;
    const auto tile_end_of = [&](int k) -> long long {
        if (k >= nchunks - 1) return tiles;
        if (k < 0) return 0;
        // TDStretch frames [0, temp + (seq_end - 1) * hop) are final; tile t reads frames up to t * stride - prefill + 2 * kFirChunks - 1
        const long long avail = (long long)temp + (long long)(seq_end_of(k) - 1) * hop;
        const long long room = avail + s->prefill - 2ll * kFirChunks;
        long long t = room < 0 ? 0 : room / kPostStride + 1;
        return t > tiles ? tiles : t;
    };
    cp->seq_begin = c == 0 ? 1 : seq_end_of(c - 1);
    cp->seq_end = seq_end_of(c);
    if (L.nseq <= 1) { cp->seq_begin = 1; cp->seq_end = 1; }
    cp->tile_begin = tile_end_of(c - 1);
    cp->tile_end = tile_end_of(c);
    if (c >= nchunks - 1) { cp->in_need = in_frames; cp->out_ready = out_frames; return; }
    // input: the search of sequence i reads its window and the region the next mid buffer comes from, the tail reads the
    // sequence itself: all below pos[i] + seek_window + seek_length
    long long need = (cp->seq_end >= 2 ? pos[(size_t)cp->seq_end - 1] : 0) + s->seek_window + s->seek_length + 8;
    cp->in_need = need < in_frames ? need : in_frames;
    // output: cubic outputs whose read position lies below tile_end * stride, i.e. i < ceil(tile_end * stride * 2^e / R)
    const unsigned __int128 num = (unsigned __int128)(cp->tile_end * (long long)kPostStride) << s->e;
    long long ready = (long long)((num + s->R - 1) / s->R);
    cp->out_ready = ready < out_frames ? ready : out_frames;
}

}  // namespace

// chunk < 0: the whole render in one go (nchunks ignored).  phase (chunked fused path): 0 = search + tail, 1 = the WSOLA
// search of the chunk only, 2 = its tail only (cross-fade + FIR + cubic over the tiles the searched sequences complete)
static int soundtouch_run_impl(nodey_soundtouch* s, float* out, int64_t out_stride, const float* in, int64_t in_stride, const TrackTab* tab,
                               int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                               int32_t* offsets, int64_t offsets_stride, int chunk, int nchunks, int phase, nodey_stream_t stream)
{
    NODEY_REQUIRE(s && out && (in || tab), NODEY_E_INVALID, "nodey_soundtouch_run: null argument");
    NODEY_REQUIRE(ntracks >= 1 && in_frames >= 0 && frame_size > 0, NODEY_E_INVALID, "nodey_soundtouch_run: bad size");
    {
        int dev = -1;
        NODEY_CUDA_OK(cudaGetDevice(&dev));
        NODEY_REQUIRE(dev == s->device, NODEY_E_INVALID, "nodey_soundtouch_run: plan belongs to device %d, current device is %d", s->device, dev);
    }
    std::lock_guard<std::mutex> lock(s->mu);
    cudaStream_t st = as_stream(stream);
    const std::shared_ptr<const nodey_soundtouch::LenPlan> lp = cached_lengths(s, in_frames, frame_size);
    StageLens L{lp->n_ext, lp->l1, lp->l2, lp->l3, lp->nseq, lp->tds_in};
    const std::vector<long long>& pos = lp->pos;
    const long long total = lp->total;
    NODEY_REQUIRE(out_frames >= 0 && out_frames <= total, NODEY_E_RANGE,
                  "nodey_soundtouch_run: out_frames %lld exceeds what SoundTouch would produce (%lld)", (long long)out_frames, total);
    if (out_frames == 0) return NODEY_OK;
    NODEY_REQUIRE(L.nseq >= 1, NODEY_E_INVALID, "nodey_soundtouch_run: internal: output without a sequence");
    NODEY_REQUIRE(offsets == nullptr || offsets_stride >= L.nseq - 1, NODEY_E_INVALID, "nodey_soundtouch_run: offsets_stride too small");
    ChunkPlan cp;
    if (chunk >= 0) {
        NODEY_REQUIRE(offsets, NODEY_E_INVALID, "nodey_soundtouch_run_chunk: the offset trace must be caller-owned (it carries the chain from chunk to chunk)");
        NODEY_REQUIRE(nchunks >= 1 && chunk < nchunks && nchunks == effective_chunks(s, L, nchunks), NODEY_E_RANGE,
                      "nodey_soundtouch_run_chunk: chunk %d of %d is not a chunking nodey_soundtouch_chunks returned", chunk, nchunks);
        chunk_plan(s, L, pos, in_frames, out_frames, chunk, nchunks, &cp);
    } else {
        cp.seq_begin = 1; cp.seq_end = (int)L.nseq; cp.tile_begin = 0; cp.tile_end = post_tiles_total(L);
        cp.in_need = in_frames; cp.out_ready = out_frames;
    }
    const bool fused = chunkable(s);
    NODEY_REQUIRE(chunk < 0 || fused || nchunks == 1, NODEY_E_RANGE, "nodey_soundtouch_run_chunk: only the fused stereo path runs in chunks");

    const long long* d_pos = nullptr;
    {
        const int rc = position_table(s, L.tds_in, pos, st, &d_pos);
        if (rc != NODEY_OK) return rc;
    }

    const int CH = s->ch;
    const long long nseq = L.nseq;
    // workspace: offsets + two intermediates per track
    const long long offs_stride_ws = nseq > 1 ? nseq - 1 : 1;
    int* d_offs = offsets;
    long long offs_stride = offsets ? offsets_stride : offs_stride_ws;
    const long long s1 = ((L.l1 * CH + 3) & ~3ll) + 4, s2 = ((L.l2 * CH + 3) & ~3ll) + 4;
    float* ws = nullptr;
    int* ws_offs = nullptr;
    if (!fused) {
        const int rc = device_alloc((void**)&ws, sizeof(float) * (size_t)((s1 + s2) * ntracks), st);
        if (rc != NODEY_OK) return rc;
    }
    if (!offsets) {
        const int rc = device_alloc((void**)&ws_offs, sizeof(int) * (size_t)(offs_stride_ws * ntracks), st);
        if (rc != NODEY_OK) { if (ws) device_free(ws, st); return rc; }
        d_offs = ws_offs;
    }
    float* b1 = ws;
    float* b2 = ws ? ws + s1 * ntracks : nullptr;

    auto run_offsets = [&](View vin, int seq_begin, int seq_end) -> int {
        TdsArgs ta;
        ta.in = vin; if (vin.use_tab) ta.tt = *tab;
        ta.pos = d_pos; ta.offs = d_offs; ta.offs_stride = offs_stride; ta.nseq = (int)nseq;
        ta.seq_begin = seq_begin; ta.seq_end = seq_end;
        ta.overlap = s->overlap; ta.seek_window = s->seek_window; ta.seek_length = s->seek_length;
        ta.Q = 4 * (CH * s->overlap / 16);
        if (seq_end > seq_begin) {
            // cluster size: spread one track over CL SMs while the batch leaves SMs idle.  Measured in the engine, where the
            // pitch and the tempo node's chains of a batch run side by side (tools/chain_trace.py, 256-thread CTAs, two per
            // SM; ms for CL = 1 / 2 / 4): 32 tracks 50.5 / 43.2 / 33.1, 48 tracks - / 46.2 / 48.6, 64 tracks 71.4 / 53.5 / 63.7,
            // 96 tracks 79.3 / 78.7 / -, 112 tracks 86.1 / 90.8 / -, 128 tracks 94.4 / 104.7 / 125.9.  So: 4 CTAs per track
            // up to 40 tracks, 2 up to about 100 (round 1, with the KT = 8 single-CTA kernel: up to 200), 1 beyond
            int CL = 1;
            if ((long long)ntracks * 8 <= kTdsCluster8Tracks * 8ll * sm_count() / 148) CL = 8;
            else if ((long long)ntracks * 4 <= 160ll * sm_count() / 148) CL = 4;
            else if ((long long)ntracks * 2 <= 208ll * sm_count() / 148) CL = 2;
            if (s->force_cluster > 0) CL = s->force_cluster;
            if (const char* env = getenv("NODEY_TDS_CLUSTER")) { const int v = atoi(env); if (v == 1 || v == 2 || v == 4 || v == 8) CL = v; }   // development override
            const int K = 4 / CH;
            const int tcount = (s->seek_length + K - 1) / K;
            // candidates per thread: a (lane, class) stream of tcount candidates is walked by whole warps, so the cost of a
            // sequence goes with warps x KT.  One CTA per track: the smallest KT that fits the stream into ONE warp
            // (456 candidates: 31 lanes x 15 instead of 57 lanes of two warps x 8 -- 6 % fewer rounded operations issued
            // and half the window loads per operation); clusters split the stream first (29 lanes x 8 resp. x 4)
            int KT = CL >= 8 ? 2 : CL >= 4 ? 4 : 8;
            if (CL == 1 && CH == 2) { const int k1 = (tcount + 30) / 31; if (k1 >= 11 && k1 <= 16) KT = k1; }   // 31 blocks + 1 (streams that start one element late)
            if (s->force_kt > 0) KT = s->force_kt;
            if (const char* env = getenv("NODEY_TDS_KT")) { const int v = atoi(env); if (v == 2 || v == 4 || v == 8 || (v >= 11 && v <= 16)) KT = v; }   // development override
            if (KT > 8 && (CH != 2 || (tcount + KT - 1) / KT / CL + ta.Q / KT + 5 > kTdsSkLarge || getenv("NODEY_TDS_RUNTIME_SK"))) KT = 8;
            const int tblocks = (tcount + KT - 1) / KT;
            ta.tb_per = (tblocks + CL - 1) / CL;
            const int ngroups = (ta.Q + KT - 1) / KT;
            int sk = ta.tb_per + ta.Q / KT + 4;
            // compile-time sub-plane strides (window loads become base + immediate); odd resp. 4 mod 8 so that the staging
            // stores of consecutive sub-planes spread over the banks
            const int sk_fixed = KT == 2 ? kTdsSkPair : KT == 4 ? 84 : KT == 8 ? 92 : kTdsSkLarge;
            bool fixed = sk <= sk_fixed && !getenv("NODEY_TDS_RUNTIME_SK");
            if (fixed) sk = sk_fixed; else if (KT > 8) sk |= 1; else while ((sk & 7) != 4) sk++;
            ta.sk = sk;
            ta.qp = ngroups * tds_ktp_rt(KT) + 8;
            const int pad_c = (K * KT) & 1 ? 2 : 1, pad_n = KT & 1 ? 2 : 1;
            ta.ncand_pad = (K * KT * ta.tb_per + pad_c * ta.tb_per + 4 + 3) & ~3;      // + pad words per K*KT candidates
            ta.npm = (KT * (ta.tb_per + 1) + pad_n * (ta.tb_per + 1) + 4 + 3) & ~3;
            const int mreg = s->seek_length + s->overlap;
            ta.vec8 = (CH == 2 && (((uintptr_t)vin.p) & 7) == 0 && (vin.stride % 2) == 0) ? 1 : 0;
            if (vin.use_tab) {
                ta.vec8 = CH == 2 ? 1 : 0;
                for (int t = 0; t < ntracks; t++) if (((uintptr_t)tab->p[2 * t]) & 7) ta.vec8 = 0;
            }
            const size_t smem = sizeof(float) * ((size_t)8 * KT * sk + (size_t)4 * ta.qp + (size_t)4 * ta.ncand_pad + (size_t)4 * ta.npm +
                                                 (size_t)((mreg * CH + 3) & ~3)) + (sizeof(double) + 5 * sizeof(int)) * (size_t)(K * KT * ta.tb_per);
            NODEY_REQUIRE(smem <= 110 * 1024, NODEY_E_RANGE, "tds_offsets: %zu bytes of shared memory per CTA exceed the two-per-SM budget", smem);
            // (KT = 2 runs four CTAs per SM when its 27 KB allow it; with a larger window it is simply less resident)
            NODEY_REQUIRE(KT * (ta.tb_per + ta.Q / KT + 4) * 4 / CH <= 2048 + 64 * KT, NODEY_E_RANGE, "tds_offsets: search window of %d frames exceeds the staged maximum", KT * sk * 4 / CH);
            void (*kern)(TdsArgs) = tds_kernel_for(CH, KT, fixed);
            NODEY_REQUIRE(kern, NODEY_E_INVALID, "tds_offsets: internal: no kernel for %d candidates per thread", KT);
            NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            NODEY_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof(cfg));
            cfg.gridDim = dim3((unsigned)(ntracks * CL)); cfg.blockDim = dim3(kTdsThreads);
            cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            {
                LaunchScope ls("tds_offsets_kernel", st);
                NODEY_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta));
            }
            NODEY_LAUNCH_OK();
        }
        return NODEY_OK;
    };
    auto run_tds = [&](View vin, float* dst, long long dst_stride, long long dst_cap) -> int {
        const int rc0 = run_offsets(vin, 1, (int)nseq);
        if (rc0 != NODEY_OK) return rc0;
        AsmArgs aa;
        aa.in = vin; if (vin.use_tab) aa.tt = *tab;
        aa.out = dst; aa.out_stride = dst_stride; aa.out_cap = dst_cap; aa.pos = d_pos;
        aa.offs = d_offs; aa.offs_stride = offs_stride; aa.fade = s->d_fade; aa.nseq = (int)nseq;
        aa.overlap = s->overlap; aa.seek_window = s->seek_window;
        dim3 grid((unsigned)nseq, (unsigned)ntracks);
        if (CH == 2) NODEY_LAUNCH("tds_assemble_kernel", st, tds_assemble_kernel<2><<<grid, 256, 0, st>>>(aa)); else NODEY_LAUNCH("tds_assemble_kernel", st, tds_assemble_kernel<1><<<grid, 256, 0, st>>>(aa));
        NODEY_LAUNCH_OK();
        return NODEY_OK;
    };
    auto run_fir = [&](View vin, float* dst, long long dst_stride, long long count) -> int {
        if (count <= 0) return NODEY_OK;
        FirArgs fa; fa.in = vin; fa.out = dst; fa.out_stride = dst_stride; fa.count = count; memcpy(fa.h, s->aa, sizeof(fa.h));
        const int tile = CH == 2 ? kFirTileS : kFirTile;
        long long tiles = (count + tile - 1) / tile;
        long long gx = tiles < 65535 ? tiles : 65535;
        dim3 grid((unsigned)gx, (unsigned)ntracks);
        const bool aligned = (((uintptr_t)dst) & 15) == 0 && (dst_stride % 4) == 0;
        NODEY_REQUIRE(CH == 1 || aligned, NODEY_E_INVALID, "aa_fir: stereo output must be 16-byte aligned");
        if (CH == 2) NODEY_LAUNCH("aa_fir_kernel", st, aa_fir_stereo_kernel<<<grid, kFirThreads, 0, st>>>(fa)); else NODEY_LAUNCH("aa_fir_kernel", st, aa_fir_mono_kernel<<<grid, kFirThreads, 0, st>>>(fa));
        NODEY_LAUNCH_OK();
        return NODEY_OK;
    };
    auto run_cubic = [&](View vin, float* dst, long long dst_stride, long long count) -> int {
        if (count <= 0) return NODEY_OK;
        CubicArgs ca; ca.in = vin; if (vin.use_tab) ca.tt = *tab;
        ca.out = dst; ca.out_stride = dst_stride; ca.count = count; ca.R = s->R; ca.e = s->e;
        long long blocks = (count + 255) / 256;
        const long long cap = (long long)sm_count() * 8;
        dim3 grid((unsigned)(blocks < cap ? blocks : cap), (unsigned)ntracks);
        if (CH == 2) NODEY_LAUNCH("cubic_kernel", st, cubic_kernel<2><<<grid, 256, 0, st>>>(ca)); else NODEY_LAUNCH("cubic_kernel", st, cubic_kernel<1><<<grid, 256, 0, st>>>(ca));
        NODEY_LAUNCH_OK();
        return NODEY_OK;
    };

    int rc = NODEY_OK;
    if (fused) {
        // offsets, then the fused assemble + FIR + cubic tail, both over this chunk's range
        View v0{in, in_frames, 0, in_stride, tab ? 1 : 0};
        if (chunk >= 0 && nchunks > 1 && chunk == 0 && phase != 2) {
            // the tail's interior fast path looks at the offsets of the next sequences before it knows whether it needs
            // them: give the not yet searched ones a defined value
            NODEY_CUDA_OK(cudaMemsetAsync(d_offs, 0, sizeof(int) * (size_t)offs_stride * (size_t)ntracks, st));
        }
        if (phase != 2) rc = run_offsets(v0, cp.seq_begin, cp.seq_end);
        if (rc == NODEY_OK && phase != 1 && cp.tile_end > cp.tile_begin) {
            PostArgs pa;
            pa.in = v0; if (v0.use_tab) pa.tt = *tab;
            pa.pos = d_pos; pa.offs = d_offs; pa.offs_stride = offs_stride; pa.fade = s->d_fade;
            pa.nseq = (int)nseq; pa.overlap = s->overlap; pa.seek_window = s->seek_window; pa.prefill = s->prefill; pa.l1 = L.l1;
            memcpy(pa.h, s->aa, sizeof(pa.h));
            pa.R = s->R; pa.e = s->e; pa.out = out; pa.out_stride = out_stride; pa.count = out_frames;
            pa.tile_outputs = (long long)((((unsigned __int128)kPostStride) << s->e) / s->R);
            pa.tile_begin = cp.tile_begin; pa.tile_end = cp.tile_end;
            pa.u.neg_zero = -0.0f; pa.u.one = 1.0f;
            const long long tiles = cp.tile_end - cp.tile_begin;
            const long long cap = (long long)sm_count() * 4;
            dim3 grid((unsigned)(tiles < cap ? tiles : cap), (unsigned)ntracks);
            NODEY_LAUNCH("st_post_kernel", st, st_post_kernel<<<grid, kFirThreads, 0, st>>>(pa));
            NODEY_LAUNCH_OK();
        }
    } else if (s->td_first) {
        View v0{in, in_frames, 0, in_stride, tab ? 1 : 0};
        rc = run_tds(v0, b1, s1, L.l1);
        View v1{b1, L.l1, s->prefill, s1, 0};
        if (rc == NODEY_OK) rc = run_fir(v1, b2, s2, L.l2);
        View v2{b2, L.l2, 0, s2, 0};
        if (rc == NODEY_OK) rc = run_cubic(v2, out, out_stride, out_frames);
    } else {
        View v0{in, in_frames, s->prefill, in_stride, tab ? 1 : 0};
        rc = run_cubic(v0, b1, s1, L.l1);
        View v1{b1, L.l1, 0, s1, 0};
        if (rc == NODEY_OK) rc = run_fir(v1, b2, s2, L.l2);
        View v2{b2, L.l2, 0, s2, 0};
        if (rc == NODEY_OK) rc = run_tds(v2, out, out_stride, out_frames);
    }
    if (ws) device_free(ws, st);
    if (ws_offs) device_free(ws_offs, st);
    return rc;
}

extern "C" {

int nodey_soundtouch_run(nodey_soundtouch* s, float* out, int64_t out_stride, const float* in, int64_t in_stride,
                         int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                         int32_t* offsets, int64_t offsets_stride, nodey_stream_t stream)
{
    return soundtouch_run_impl(s, out, out_stride, in, in_stride, nullptr, ntracks, in_frames, frame_size, out_frames, offsets, offsets_stride, -1, 1, 0, stream);
}

static int make_tab(const nodey_soundtouch* s, TrackTab* tab, const float* const* in_a, const float* const* in_b, int ntracks)
{
    NODEY_REQUIRE(s && in_a, NODEY_E_INVALID, "nodey_soundtouch_run_tracks: null argument");
    NODEY_REQUIRE(ntracks >= 1 && ntracks <= kMaxTabTracks, NODEY_E_RANGE, "nodey_soundtouch_run_tracks: 1..%d tracks per call", kMaxTabTracks);
    memset(tab, 0, sizeof(*tab));
    for (int t = 0; t < ntracks; t++) {
        NODEY_REQUIRE(in_a[t], NODEY_E_INVALID, "nodey_soundtouch_run_tracks: null track pointer");
        tab->p[2 * t] = in_a[t];
        tab->p[2 * t + 1] = (in_b && s->ch == 2) ? in_b[t] : nullptr;
    }
    return NODEY_OK;
}

int nodey_soundtouch_run_tracks(nodey_soundtouch* s, float* out, int64_t out_stride, const float* const* in_a, const float* const* in_b,
                                int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                                int32_t* offsets, int64_t offsets_stride, nodey_stream_t stream)
{
    TrackTab tab;
    const int rc = make_tab(s, &tab, in_a, in_b, ntracks);
    if (rc != NODEY_OK) return rc;
    return soundtouch_run_impl(s, out, out_stride, nullptr, 0, &tab, ntracks, in_frames, frame_size, out_frames, offsets, offsets_stride, -1, 1, 0, stream);
}

/* SURVEY.md App. C7.  What soundtouch_process_payload (audio-velocity.cpp:286-441) emits when an input frame is available
 * at every turn of its loop: one putSamples per turn, then receiveSamples(min(numSamples, 3 * 1152 / velocity)) whenever more
 * than 1152 / velocity samples are queued; at the end of the input the loop leaves through `if (numSamples() == 0 &&
 * input_stream_eof) break;` (:414) BEFORE it ever reaches flush() whenever the last receive emptied the FIFO -- which is the
 * normal case -- and what SoundTouch still holds is lost.  Host arithmetic on lengths only: numSamples() after k frames is
 * the unflushed output length of the first k frames minus what was received.  Returns the total the node emits (a prefix of
 * the canonical, always-flushed render) and its frame sizes, run-length encoded. */
int64_t nodey_soundtouch_reference_schedule(nodey_soundtouch* s, int64_t in_frames, int frame_size, float velocity,
                                            int64_t* run_len, int64_t* run_count, int64_t run_cap, int64_t* n_runs, int* flushed)
{
    if (!s || in_frames < 0 || frame_size <= 0 || !(velocity > 0.f)) return NODEY_E_INVALID;
    std::lock_guard<std::mutex> lock(s->mu);
    // sequence positions are prefix stable in the input length: build them once for the longest input flush() can make
    std::vector<long long> pos;
    {
        StageLens Lmax;
        stage_lengths(s, in_frames + 128ll * 200, &Lmax, &pos);
    }
    const auto out_len = [&](long long n_ext) -> long long {
        // stage_lengths() with the sequence count read off the table: sequences whose window fits into the TDStretch input
        const auto nseq_of = [&](long long tds_in) -> long long {
            long long lo = 0, hi = (long long)pos.size();
            while (lo < hi) { const long long mid = (lo + hi) / 2; if (pos[(size_t)mid] + s->sample_req <= tds_in) lo = mid + 1; else hi = mid; }
            return lo;
        };
        if (s->td_first) return cubic_count(s, fir_count(s, s->prefill + tds_out_frames(s, nseq_of(n_ext))));
        return tds_out_frames(s, nseq_of(fir_count(s, cubic_count(s, s->prefill + n_ext))));
    };
    const double time_ratio = 1.0f / velocity;
    const unsigned min_samples = (unsigned)(time_ratio * 1152), max_samples = (unsigned)(time_ratio * 1152 * 3);
    const double denom = s->rate * s->tempo;
    double expected = 0;
    long long fed = 0, received = 0;
    bool eof = false, created = false, did_flush = false;
    std::vector<std::pair<long long, long long>> sizes;       // frame sizes the node pushes, run-length encoded
    const auto emit = [&](long long n) {
        if (!sizes.empty() && sizes.back().first == n) sizes.back().second++;
        else sizes.emplace_back(n, 1);
    };
    for (;;) {
        if (!eof) {
            if (fed >= in_frames) eof = true;
            else {
                const long long n = (in_frames - fed) < frame_size ? (in_frames - fed) : frame_size;
                fed += n; expected += (double)n / denom; created = true;
            }
        }
        if (!created) { if (eof) break; continue; }
        long long avail = out_len(fed) - received;
        if (avail == 0 && eof) break;
        if ((unsigned long long)avail > min_samples) {
            const long long take = avail < (long long)max_samples ? avail : (long long)max_samples;
            received += take; emit(take);
        } else if (eof) {
            int still = (int)((long)(expected + 0.5) - (long)received);
            if (still < 0) still = 0;
            long long ext = fed;
            for (int i = 0; still > (int)(out_len(ext) - received) && i < 200; i++) ext += 128;
            avail = out_len(ext) - received;
            if (avail > still) avail = still;
            did_flush = true;
            if (avail > 0) { received += avail; emit(avail); }
            break;
        }
    }
    for (size_t k = 0; k < sizes.size() && (long long)k < run_cap && run_len && run_count; k++) { run_len[k] = sizes[k].first; run_count[k] = sizes[k].second; }
    if (n_runs) *n_runs = (long long)sizes.size();
    if (flushed) *flushed = did_flush ? 1 : 0;
    return received;
}

int nodey_soundtouch_chunks(nodey_soundtouch* s, int64_t in_frames, int frame_size, int64_t out_frames, int want_chunks,
                            int64_t* in_need, int64_t* out_ready, int cap)
{
    NODEY_REQUIRE(s && in_frames >= 0 && frame_size > 0 && out_frames >= 0, NODEY_E_INVALID, "nodey_soundtouch_chunks: bad argument");
    std::lock_guard<std::mutex> lock(s->mu);
    const std::shared_ptr<const nodey_soundtouch::LenPlan> lp = cached_lengths(s, in_frames, frame_size);
    StageLens L{lp->n_ext, lp->l1, lp->l2, lp->l3, lp->nseq, lp->tds_in};
    const std::vector<long long>& pos = lp->pos;
    const long long total = lp->total;
    NODEY_REQUIRE(out_frames <= total, NODEY_E_RANGE, "nodey_soundtouch_chunks: out_frames %lld exceeds what SoundTouch would produce (%lld)",
                  (long long)out_frames, total);
    const int n = out_frames == 0 ? 1 : effective_chunks(s, L, want_chunks);
    for (int c = 0; c < n && c < cap; c++) {
        ChunkPlan cp;
        if (out_frames == 0) { cp.in_need = in_frames; cp.out_ready = 0; }
        else chunk_plan(s, L, pos, in_frames, out_frames, c, n, &cp);
        if (in_need) in_need[c] = cp.in_need;
        if (out_ready) out_ready[c] = cp.out_ready;
    }
    return n;
}

int nodey_soundtouch_run_chunk(nodey_soundtouch* s, float* out, int64_t out_stride, const float* in, int64_t in_stride,
                               int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                               int32_t* offsets, int64_t offsets_stride, int chunk, int nchunks, int phase, nodey_stream_t stream)
{
    NODEY_REQUIRE(chunk >= 0 && phase >= 0 && phase <= 2, NODEY_E_INVALID, "nodey_soundtouch_run_chunk: bad chunk or phase");
    return soundtouch_run_impl(s, out, out_stride, in, in_stride, nullptr, ntracks, in_frames, frame_size, out_frames, offsets, offsets_stride, chunk, nchunks, phase, stream);
}

int nodey_soundtouch_run_tracks_chunk(nodey_soundtouch* s, float* out, int64_t out_stride, const float* const* in_a, const float* const* in_b,
                                      int ntracks, int64_t in_frames, int frame_size, int64_t out_frames,
                                      int32_t* offsets, int64_t offsets_stride, int chunk, int nchunks, int phase, nodey_stream_t stream)
{
    NODEY_REQUIRE(chunk >= 0 && phase >= 0 && phase <= 2, NODEY_E_INVALID, "nodey_soundtouch_run_tracks_chunk: bad chunk or phase");
    TrackTab tab;
    const int rc = make_tab(s, &tab, in_a, in_b, ntracks);
    if (rc != NODEY_OK) return rc;
    return soundtouch_run_impl(s, out, out_stride, nullptr, 0, &tab, ntracks, in_frames, frame_size, out_frames, offsets, offsets_stride, chunk, nchunks, phase, stream);
}

}  // extern "C"
