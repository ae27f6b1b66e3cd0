// Internal interface between typing.cu and em.cu: batched EM on class tables that already live on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct hgt_ctx;

struct EmDevProblem {
    const uint64_t *bits;            // [C_max][wp] class rows (contiguous region of a class pool)
    const unsigned long long *cnt;   // [C_max] class counts
    const int32_t *class_first;      // [C_max] first pair index of each class (tie-break key)
    const int32_t *C_ptr;            // device counter: number of classes actually present (<= C_max)
    int C_max, A, wp;
    int A_live_max;                  // upper bound of alleles that can be members of a class (table / keep mask size)
    const double *len;               // [A] or null
    int remove_low;
    double *prob;                    // [A]
    uint8_t *in_result;              // [A]
    int32_t *first_class;            // [A]
    int32_t *iters_status;           // [3]
    void *ws;                        // hgt_em_problem_ws_bytes(wp) bytes of device scratch
    void *coop_ws;                   // non-null: run this problem on all SMs (cooperative launch); needs
                                     // hgt_em_coop_ws_bytes(ctx, A) bytes
};

size_t hgt_em_problem_ws_bytes(int wp);
size_t hgt_em_args_bytes(int n_problems);
// true when a problem of this shape is too large for one SM's shared memory and big enough to deserve the whole GPU - or
// when the launch holds only a handful of problems (n_problems), so that one streaming CTA per problem would leave the
// GPU idle (a single sample of 10 k single-end reads on a 7 k-allele locus: 15.5 ms on one SM)
bool hgt_em_wants_coop(const hgt_ctx *ctx, int C, int A, int wp, int A_live_max, int n_problems);
size_t hgt_em_coop_ws_bytes(const hgt_ctx *ctx, int A);
// One CTA per problem, all problems of all loci in as few launches as the register variants need.  h_args: host
// staging (pinned) and d_args: device copy, hgt_em_args_bytes(n_problems) each; h_args must stay untouched until the
// stream has passed this call.
int hgt_em_batch_dev(hgt_ctx *ctx, cudaStream_t st, int n_problems, const EmDevProblem *problems, void *h_args,
                     void *d_args);
