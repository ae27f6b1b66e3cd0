// Internal interface between typing.cu and em.cu: batched EM on class tables that already live on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct hgt_ctx;

struct EmDevProblem {
    const uint64_t *bits;            // [C_max][wp] class rows (contiguous region of a class pool)
    const unsigned long long *cnt;   // [C_max] class counts
    const int32_t *class_first;      // [C_max] first pair index of each class (tie-break key)
    const int32_t *C_ptr;            // device counter: number of classes actually present
    int C_max, A;
    const double *len;               // [A] or null
    int remove_low;
    double *prob;                    // [A]
    uint8_t *in_result;              // [A]
    int32_t *first_class;            // [A]
    int32_t *iters_status;           // [3]
};

size_t hgt_em_batch_ws_bytes(int n_problems, int wp);
int hgt_em_batch_dev(hgt_ctx *ctx, cudaStream_t st, int n_problems, const EmDevProblem *problems, int wp, void *ws);
