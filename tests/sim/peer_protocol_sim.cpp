// Host-thread model of the peer-memory exchange protocol of picturedrocks_b200/csrc/peer.cu
// (TEST INFRASTRUCTURE: it restates the protocol — slots, flags, step numbers, the order
// gather -> commit -> clean -> scatter -> column — with one thread per rank and
// std::atomic release/acquire in place of st.release.sys / ld.acquire.sys, and injects random
// delays between the phases to shake out ordering bugs: a record overwritten before it was
// read, a column cleaned while a peer still pulls it, a flag skipped).
//
//   peer_protocol_sim <world> <cells> <steps> <seed> [mutation]  -> "ok" or the first violation
// mutation (to show that the model has teeth): 1 = candidate slots not double-buffered,
// 2 = the x_j area is cleaned BEFORE the step's gather instead of after it.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

struct Rec {
    double score;
    int64_t idx;
};

struct Buf {  // one rank's symmetric buffer
    std::vector<Rec> cand;                       // [2][W]
    std::vector<std::atomic<uint32_t>> flag;     // [2][W]
    std::atomic<uint32_t> cflag{0};
    std::vector<uint8_t> col;                    // [N]
    Buf(int W, int N) : cand(2 * W), flag(2 * W), col(N, 0)
    {
        for (auto &f : flag) f.store(0);
    }
};

static int W, N, STEPS, GENES_PER_RANK = 5, MUT = 0;
static std::vector<Buf *> bufs;
static std::atomic<int> failed{0};

static uint8_t code_of(int gene, int cell) { return (uint8_t)(((gene * 2654435761u) ^ (cell * 40503u)) % 5u); }
static double score_of(int gene, int step) { return (double)(((gene + 1) * 7919u + step * 104729u) % 1000u); }

static void jitter(std::mt19937 &rng)
{
    const int r = rng() % 8;
    if (r == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 300));
    else if (r < 4) std::this_thread::yield();
}

static bool timed_out(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::steady_clock::now() - t0 > std::chrono::seconds(6);
}

static void fail(const char *what, int rank, int step)
{
    if (!failed.exchange(1)) std::printf("VIOLATION: %s (rank %d, step %d)\n", what, rank, step);
}

static void rank_main(int rank, unsigned seed)
{
    std::mt19937 rng(seed * 977u + rank);
    Buf &mine = *bufs[rank];
    uint32_t seq = 0;
    int dirty = 0;
    std::vector<uint8_t> xj(N);
    for (int s = 0; s < STEPS && !failed.load(); ++s) {
        if (MUT == 2 && dirty) std::fill(mine.col.begin(), mine.col.end(), 0);
        // ---- gather16: local best, push to every peer, wait for all records -------------
        int best_g = rank * GENES_PER_RANK;
        for (int g = rank * GENES_PER_RANK; g < (rank + 1) * GENES_PER_RANK; ++g)
            if (score_of(g, s) > score_of(best_g, s)) best_g = g;
        const uint32_t step = seq + 1;
        const int par = MUT == 1 ? 0 : (step & 1);
        for (int p = 0; p < W; ++p) {
            jitter(rng);
            bufs[p]->cand[par * W + rank] = Rec{score_of(best_g, s), best_g};
            bufs[p]->flag[par * W + rank].store(step, std::memory_order_release);
        }
        std::vector<Rec> out(W);
        for (int p = 0; p < W; ++p) {
            const auto t0 = std::chrono::steady_clock::now();
            while (mine.flag[par * W + p].load(std::memory_order_acquire) != step) {
                if (failed.load()) return;
                if (timed_out(t0)) return fail("candidate flag never showed this step (overwritten?)", rank, s);
                std::this_thread::yield();
            }
            out[p] = mine.cand[par * W + p];
        }
        seq = step;
        jitter(rng);
        // ---- commit: identical reduce on every rank ---------------------------------------
        int win = 0;
        for (int p = 1; p < W; ++p)
            if (out[p].score > out[win].score || (out[p].score == out[win].score && out[p].idx < out[win].idx))
                win = p;
        const int gene = (int)out[win].idx;
        // every record must be the one its sender computed for THIS step
        for (int p = 0; p < W; ++p) {
            int bg = p * GENES_PER_RANK;
            for (int g = p * GENES_PER_RANK; g < (p + 1) * GENES_PER_RANK; ++g)
                if (score_of(g, s) > score_of(bg, s)) bg = g;
            if (out[p].idx != bg || out[p].score != score_of(bg, s)) fail("stale or torn candidate record", rank, s);
        }
        const int owner = gene / GENES_PER_RANK;
        // ---- clean (after this step's gather), scatter (owner) --------------------------------
        jitter(rng);
        if (MUT != 2 && dirty) std::fill(mine.col.begin(), mine.col.end(), 0);
        jitter(rng);
        if (owner == rank)
            for (int c = 0; c < N; ++c) mine.col[c] = code_of(gene, c);
        // ---- column: owner publishes, the others wait and pull ----------------------------------
        jitter(rng);
        if (owner == rank) {
            for (int p = 0; p < W; ++p) bufs[p]->cflag.store(seq, std::memory_order_release);
        } else {
            const auto t0 = std::chrono::steady_clock::now();
            while (mine.cflag.load(std::memory_order_acquire) != seq) {
                if (failed.load()) return;
                if (timed_out(t0)) return fail("column flag never showed this step (overwritten?)", rank, s);
                std::this_thread::yield();
            }
        }
        for (int c = 0; c < N; ++c) {
            if ((c & 63) == 0) jitter(rng);
            xj[c] = bufs[owner]->col[c];
        }
        dirty = owner == rank;
        for (int c = 0; c < N; ++c)
            if (xj[c] != code_of(gene, c)) {
                fail("pulled column differs from the winner's column", rank, s);
                break;
            }
        // non-owners' own areas must still be all-zero (nobody else writes them)
        if (owner != rank)
            for (int c = 0; c < N; ++c)
                if (mine.col[c] != 0) {
                    fail("a non-owner's x_j area is not clean", rank, s);
                    break;
                }
        jitter(rng);  // the rest of the greedy step (stream, finalize, update)
    }
}

int main(int argc, char **argv)
{
    W = argc > 1 ? std::atoi(argv[1]) : 4;
    N = argc > 2 ? std::atoi(argv[2]) : 512;
    STEPS = argc > 3 ? std::atoi(argv[3]) : 200;
    const unsigned seed = argc > 4 ? (unsigned)std::atoi(argv[4]) : 1u;
    MUT = argc > 5 ? std::atoi(argv[5]) : 0;
    for (int r = 0; r < W; ++r) bufs.push_back(new Buf(W, N));
    std::vector<std::thread> th;
    for (int r = 0; r < W; ++r) th.emplace_back(rank_main, r, seed);
    for (auto &t : th) t.join();
    if (!failed.load()) std::printf("ok\n");
    return failed.load();
}
