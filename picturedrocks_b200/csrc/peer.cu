// peer.cu — the per-step exchange of the gene-sharded greedy loop over NVLink peer memory.
//
// The reference takes np.argmax over the whole score vector (picturedrocks/markers/
// _iterative.py:64-66).  With the genes sharded over the GPUs of one NVSwitch box every rank
// holds one (score, global index) candidate per step and all ranks need all of them: 16 bytes
// per rank.  An NCCL all-gather of that size costs ~18 us of pure latency per greedy step (a
// tenth of the step on 8 GPUs); here every rank STORES its record straight into a slot of
// every peer's buffer (the buffers are mapped into each other's address space through CUDA
// IPC, see dist.PeerRecords) and the next step's prb_pick_scatter waits for the slots.
//
// Buffer of a rank: uint64 slot[2][W][4] = {score bits, index, tag, unused}; the parity of the
// tag picks the half, so a rank that runs one step ahead never overwrites a record a slower
// peer has not read yet (it cannot run two steps ahead: it needs the peer's next record first —
// prb_pick_scatter waits for the W tags at every step that is followed by a send, including
// the add / remove steps that use no record).
// tag = number of exchanges so far (a device counter that only this kernel advances; it is the
// same on all ranks because they all run the same sequence of steps).  The record is written
// first, then the tag with a release store at system scope; readers acquire the tag.
#include "select.cuh"

namespace prb {

__global__ void __launch_bounds__(32)
k_peer_share(const double *__restrict__ best, unsigned long long *const *__restrict__ peer_bufs,
             int W, int rank, int *xtag)
{
    const int tag = *xtag + 1;
    __syncwarp();
    for (int r = threadIdx.x; r < W; r += 32) {
        unsigned long long *slot = peer_bufs[r] + ((size_t)(tag & 1) * W + rank) * 4;
        const unsigned long long sc = (unsigned long long)__double_as_longlong(best[0]);
        const unsigned long long ix = ((const unsigned long long *)best)[1];
        asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(slot), "l"(sc), "l"(ix) : "memory");
        __threadfence_system();
        asm volatile("st.release.sys.global.b64 [%0], %1;" ::"l"(slot + 2),
                     "l"((unsigned long long)tag)
                     : "memory");
    }
    __syncwarp();
    if (threadIdx.x == 0) *xtag = tag;
}

}  // namespace prb

using namespace prb;

extern "C" int prb_peer_share(const double *best, const uint64_t *const *peer_bufs, int W, int rank,
                              int32_t *xtag, prb_stream_t stream)
{
    PRB_REQUIRE(best && peer_bufs && xtag && W >= 1 && W <= 1024 && rank >= 0 && rank < W,
                "bad arguments");
    k_peer_share<<<1, 32, 0, S(stream)>>>(best, (unsigned long long *const *)peer_bufs, W, rank, xtag);
    PRB_LAUNCH_CHECK("k_peer_share");
    return 0;
}
