// Host-side launchers of the engine's kernels (internal; the public surface is
// include/maxpro_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "common.cuh"

namespace mp {

constexpr size_t kSmemBudget = 227 * 1024;  // usable dynamic shared memory per CTA on sm_100
constexpr int kSmallMaxD = 8;               // thread-per-candidate kernels are instantiated for d <= 8

// ---- K1+K2: generate-and-score search ---------------------------------------
struct SearchLaunch {
    unsigned long long n, d;
    int metric;  // 0 maxpro, 1 maximin
    unsigned long long seed, lo, hi;
    uint32_t flags;
    BestRec* block_best;         // [grid]
    unsigned int* done_counter;  // zero before first use; the kernel re-arms it
    BestRec* result;
    int grid, threads;
    int group;  // lanes per candidate (1 or 2)
    const unsigned long long* list;  // optional explicit candidate list (cyclic kernel, stage 2 of the screened search)
    unsigned long long list_count;
    const double* designs;           // supplied mode (cyclic kernel): score designs[lo..hi) instead of generating
    double* scores;
    cudaStream_t stream;
};
bool search_small_supported(unsigned long long n, unsigned long long d);
void search_small_pick_config(unsigned long long n, unsigned long long d, int* group, int* threads);
cudaError_t launch_search_small(const SearchLaunch& L);

// shape-specialised cyclic schedule (search_cyclic.cu); *cands = candidates per CTA
bool search_cyclic_supported(unsigned long long n, unsigned long long d, int* cands);
cudaError_t launch_search_cyclic(const SearchLaunch& L);

// fp32 screening pass (search_screen.cu): one surviving candidate index per warp
bool search_screen_supported(unsigned long long n, unsigned long long d, int* cands);
cudaError_t launch_search_screen(const SearchLaunch& L, unsigned long long* survivors);

// warp-per-candidate kernel for medium designs (search_warp.cu): 16 candidates per SM in shared memory
bool search_warp_supported(unsigned long long n, unsigned long long d);
int search_warp_grid(unsigned long long n, unsigned long long d, int sms);
cudaError_t launch_search_warp(const SearchLaunch& L);

// CTA-per-candidate kernel for designs that do not fit the thread-group kernels (search_cta.cu)
bool search_cta_supported(unsigned long long n, unsigned long long d);
int search_cta_grid(unsigned long long n, unsigned long long d, int sms);
cudaError_t launch_search_cta(const SearchLaunch& L);

// ---- K1 alone: materialise designs (winner regeneration, generate_lhd) -------
// out: count designs, row-major n*d each; design c uses stream seed + lo + c.
cudaError_t launch_generate(unsigned long long n, unsigned long long d, unsigned long long seed,
                            unsigned long long lo, unsigned long long count, double* d_out,
                            cudaStream_t stream);

// ---- K2 on supplied designs ---------------------------------------------------
struct ScoreLaunch {
    const double* designs;  // [b][n][d] device
    unsigned long long b, n, d;
    int metric;
    double* scores;          // [b] device (required)
    double* partial;         // [b * items] device scratch
    BestRec* block_best;     // [finalize grid]
    unsigned int* done_counter;
    BestRec* result;         // may be null
    cudaStream_t stream;
};
unsigned long long score_items(unsigned long long n);  // tile pairs per design
int score_finalize_grid(unsigned long long b);
cudaError_t launch_score_batch(const ScoreLaunch& L);

// ---- K3: annealing chains ---------------------------------------------------------
struct AnnealLaunch {
    const double* design;  // device, [n][d]
    unsigned long long n, d, iters, seed, n_chains;
    double t0, cooling;
    int metric, minimize, swap;
    double* best_design;     // [n_chains][n*d]
    double* best_metric;     // [n_chains]
    unsigned int* improved;  // [n_chains]
    double* raw0;            // device scratch, 1 double: S (or min squared distance) of the start design
    double* cur_scratch;     // device scratch for chains whose design (or its proposal) does not fit shared
                             // memory: anneal_jitter_scratch_bytes(n, d, swap, n_chains) bytes, null when that is 0
    double* rowmin0;         // device scratch, n doubles and n 16-bit indices: nearest-neighbour table of the
    unsigned short* rowarg0; // start design (incremental maximin chains, anneal_maximin.cu)
    cudaStream_t stream;
};
bool anneal_supported(unsigned long long n, unsigned long long d, int swap);
size_t anneal_jitter_scratch_bytes(unsigned long long n, unsigned long long d, int swap, unsigned long long n_chains);  // 0: not needed
cudaError_t launch_anneal(const AnnealLaunch& L);
// swap moves on MaxPro as a two-stage pipeline (anneal_pipe.cu); launch_anneal dispatches to it
bool anneal_pipe_supported(unsigned long long n, unsigned long long d, int metric, int swap);
cudaError_t launch_anneal_pipe(const AnnealLaunch& L);
// swap moves on maximin with an exact nearest-neighbour table (anneal_maximin.cu); launch_anneal dispatches to it
bool anneal_maximin_supported(unsigned long long n, unsigned long long d, int metric, int swap);
cudaError_t launch_anneal_maximin(const AnnealLaunch& L);

// ---- K4: greedy ordering ------------------------------------------------------------
struct OrderLaunch {
    const double* design;     // device, [n][d]
    unsigned long long n, d;
    int metric, minimize;
    unsigned long long* perm; // device, [n]
    double* acc;              // device scratch, [n]
    unsigned int* pool;       // device scratch, [n]
    cudaStream_t stream;
};
cudaError_t launch_order(const OrderLaunch& L);
// cluster / distributed-shared-memory version (order_cluster.cu): pool spread over 8 SMs
bool order_cluster_supported(unsigned long long n, unsigned long long d);
cudaError_t launch_order_cluster(const OrderLaunch& L);

// ---- FP64 pipe peak probe -------------------------------------------------------
cudaError_t launch_dfma_probe(double* d_sink, int grid, int threads, int iters, cudaStream_t stream);

void count_launch(int k = 1);

}  // namespace mp
