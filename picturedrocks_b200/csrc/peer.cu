// peer.cu — the per-step exchange of the gene-sharded selection over NVLink peer memory
// (EXPERIMENTAL, off by default: PRB_PEER_EXCHANGE=1; written in round 1, first hardware run
// planned for round 2 — see DESIGN.md §5).
//
// The two exchanges of a greedy step (picturedrocks_b200/dist.py) as plain kernels on a
// SYMMETRIC buffer (same layout on every rank, every rank holds the device pointers of all
// peers' buffers; torch.distributed._symmetric_memory provides the allocation):
//
//   1. k_peer_gather16: local argmax of the block candidates -> the 16-byte record
//      (score, global index) is STORED into slot [parity][rank] of every peer, followed by a
//      release store of the step number into the peer's flag [parity][rank]; then the
//      kernel waits until all flags of its own buffer carry the step number and copies the
//      W records out (the unchanged prb_argmax_commit reduces them on every rank).
//      Replaces prb_argmax_final + the NCCL all-gather.  Slots are double-buffered by the
//      parity of the step number: a rank can run at most one step ahead of a peer (its next
//      gather waits for that peer's record).
//   2. The owner of the winning gene scatters its column into the x_j area of ITS symmetric
//      buffer (prb_scatter_gene_u8, unchanged), then k_peer_column: the owner publishes the
//      step number in every peer's column flag and copies the column into its local x_j;
//      every other rank waits for the flag and PULLS the column out of the owner's buffer
//      (16-byte loads over NVLink) into its local x_j.  Replaces the max-all-reduce.
//      The owner's x_j area is zeroed again by k_peer_clean of its NEXT step, after that
//      step's gather — which proves that every peer has finished pulling.
//
// Memory order: data stores, __threadfence_system(), st.release.sys flag  |  ld.acquire.sys
// flag, data loads.  A wait that lasts longer than ~10 s traps (a lost peer must not hang
// the GPU).
#include "common.cuh"

namespace prb {

constexpr int PEER_BLOCK = 256;
constexpr int PEER_MAX_WORLD = 64;

struct PeerLayout {
    int64_t off_cand;   // [2][W] records of 16 bytes
    int64_t off_flag;   // [2][W] uint32 step numbers
    int64_t off_cflag;  // uint32 step number of the published column
    int64_t off_col;    // uint8 [n_pad] dense codes of the winning gene
    int64_t n_pad;      // N rounded up to 16
    int64_t bytes;
};

__host__ __device__ inline PeerLayout peer_layout(int64_t N, int W)
{
    PeerLayout L;
    L.off_cand = 0;
    L.off_flag = (int64_t)2 * W * 16;
    L.off_cflag = (L.off_flag + (int64_t)2 * W * 4 + 127) & ~(int64_t)127;
    L.off_col = L.off_cflag + 128;
    L.n_pad = (N + 15) & ~(int64_t)15;
    L.bytes = L.off_col + L.n_pad;
    return L;
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// spin until *p == want (step numbers only grow; a newer one cannot appear before this
// rank has sent its own next record)
__device__ __forceinline__ void wait_flag(const uint32_t *p, uint32_t want)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(p) != want) {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000LL) __trap();  // ~10 s at 2 GHz
    }
}

// One block.  peers: device array of the W symmetric-buffer base pointers (rank order).
// seq: device uint32 step counter of this rank (incremented here).  out: [W][2] doubles
// (score, index bits), the layout of Comm.all_gather_cat(best).
__global__ void __launch_bounds__(PEER_BLOCK)
k_peer_gather16(const double *__restrict__ cand_score, const int64_t *__restrict__ cand_idx,
                int64_t n_cand, const uint64_t *__restrict__ peers, int rank, int W, int64_t N,
                uint32_t *seq, double *__restrict__ out)
{
    __shared__ double ss[PEER_BLOCK / 32];
    __shared__ int64_t si[PEER_BLOCK / 32];
    __shared__ double s_rec[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // local argmax of the block candidates (np.argmax order: see better())
    double sc = 0.0;
    int64_t idx = -1;
    for (int64_t c = tid; c < n_cand; c += PEER_BLOCK) {
        const double s2 = cand_score[c];
        const int64_t i2 = cand_idx[c];
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double s2 = __shfl_down_sync(0xffffffffu, sc, o);
        const int64_t i2 = __shfl_down_sync(0xffffffffu, idx, o);
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    if (lane == 0) {
        ss[warp] = sc;
        si[warp] = idx;
    }
    __syncthreads();
    if (tid == 0) {
        sc = ss[0];
        idx = si[0];
        for (int w = 1; w < PEER_BLOCK / 32; ++w)
            if (si[w] >= 0 && (idx < 0 || better(ss[w], si[w], sc, idx))) {
                sc = ss[w];
                idx = si[w];
            }
        s_rec[0] = sc;
        s_rec[1] = __longlong_as_double(idx);
    }
    __syncthreads();
    const uint32_t step = *seq + 1u;  // every thread reads the old value; written below
    const int par = (int)(step & 1u);
    const PeerLayout L = peer_layout(N, W);
    if (tid < W) {
        // push the record into slot [par][rank] of peer `tid`, then publish it
        unsigned char *base = (unsigned char *)peers[tid];
        double2 *slot = (double2 *)(base + L.off_cand) + (par * W + rank);
        *slot = make_double2(s_rec[0], s_rec[1]);
        __threadfence_system();
        st_release_sys((uint32_t *)(base + L.off_flag) + (par * W + rank), step);
    }
    __syncthreads();
    unsigned char *mine = (unsigned char *)peers[rank];
    if (tid < W) {
        wait_flag((const uint32_t *)(mine + L.off_flag) + (par * W + tid), step);
        const volatile double *slot = (const volatile double *)(mine + L.off_cand) + 2 * (par * W + tid);
        out[2 * tid] = slot[0];
        out[2 * tid + 1] = slot[1];
    }
    __syncthreads();
    if (tid == 0) *seq = step;
}

// Zero the x_j area of this rank's symmetric buffer if the previous step left this rank's
// column in it (*dirty).  Launched after the step's gather.
__global__ void k_peer_clean(unsigned char *mine, int64_t N, int W, const int32_t *dirty)
{
    if (*dirty == 0) return;
    const PeerLayout L = peer_layout(N, W);
    uint4 *col = (uint4 *)(mine + L.off_col);
    const int64_t n16 = L.n_pad >> 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16;
         i += (int64_t)gridDim.x * blockDim.x)
        col[i] = make_uint4(0, 0, 0, 0);
}

// After the owner's scatter into its symmetric x_j area.  shard_lo int32[W+1]: gene ranges.
// xj_local: uint8[n_pad] of this rank (what prb_state_v reads and zeroes).
__global__ void __launch_bounds__(PEER_BLOCK)
k_peer_column(const uint64_t *__restrict__ peers, int rank, int W, int64_t N,
              const int32_t *__restrict__ gene_ptr, const int32_t *__restrict__ shard_lo,
              const uint32_t *__restrict__ seq, unsigned char *__restrict__ xj_local,
              int32_t *dirty)
{
    const PeerLayout L = peer_layout(N, W);
    const int g = *gene_ptr;
    int owner = 0;
    while (owner + 1 < W && g >= shard_lo[owner + 1]) ++owner;
    const uint32_t step = *seq;
    unsigned char *mine = (unsigned char *)peers[rank];
    if (owner == rank) {
        if (blockIdx.x == 0 && threadIdx.x < W) {
            // the scatter kernel has completed (stream order): publish the column
            __threadfence_system();
            st_release_sys((uint32_t *)((unsigned char *)peers[threadIdx.x] + L.off_cflag), step);
        }
    } else {
        if (threadIdx.x == 0) wait_flag((const uint32_t *)(mine + L.off_cflag), step);
        __syncthreads();
    }
    const uint4 *src = (const uint4 *)((const unsigned char *)peers[owner] + L.off_col);
    uint4 *dst = (uint4 *)xj_local;
    const int64_t n16 = L.n_pad >> 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16;
         i += (int64_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(src + i)
                     : "memory");
        dst[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *dirty = owner == rank ? 1 : 0;
}

}  // namespace prb

using namespace prb;

extern "C" int prb_peer_layout(int64_t N, int W, int64_t *bytes, int64_t *off_col, int64_t *n_pad)
{
    PRB_REQUIRE(N >= 1 && W >= 1 && W <= PEER_MAX_WORLD, "bad arguments");
    const PeerLayout L = peer_layout(N, W);
    if (bytes) *bytes = L.bytes;
    if (off_col) *off_col = L.off_col;
    if (n_pad) *n_pad = L.n_pad;
    return 0;
}

extern "C" int prb_peer_gather16(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                                 const uint64_t *peers, int rank, int W, int64_t N, uint32_t *seq,
                                 double *out, prb_stream_t stream)
{
    PRB_REQUIRE(W >= 1 && W <= PEER_MAX_WORLD && rank >= 0 && rank < W, "bad rank / world size");
    k_peer_gather16<<<1, PEER_BLOCK, 0, S(stream)>>>(cand_score, cand_idx, n_cand, peers, rank, W, N,
                                                     seq, out);
    PRB_LAUNCH_CHECK("k_peer_gather16");
    return 0;
}

extern "C" int prb_peer_clean(void *mine, int64_t N, int W, const int32_t *dirty,
                              prb_stream_t stream)
{
    k_peer_clean<<<devinfo().sm_count, PEER_BLOCK, 0, S(stream)>>>((unsigned char *)mine, N, W, dirty);
    PRB_LAUNCH_CHECK("k_peer_clean");
    return 0;
}

extern "C" int prb_peer_column(const uint64_t *peers, int rank, int W, int64_t N,
                               const int32_t *gene_ptr, const int32_t *shard_lo,
                               const uint32_t *seq, uint8_t *xj_local, int32_t *dirty,
                               prb_stream_t stream)
{
    PRB_REQUIRE(W >= 1 && W <= PEER_MAX_WORLD && rank >= 0 && rank < W, "bad rank / world size");
    k_peer_column<<<devinfo().sm_count, PEER_BLOCK, 0, S(stream)>>>(peers, rank, W, N, gene_ptr,
                                                                   shard_lo, seq, xj_local, dirty);
    PRB_LAUNCH_CHECK("k_peer_column");
    return 0;
}
