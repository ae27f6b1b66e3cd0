// Diploid allele-pair model: the EM over allele PAIRS of the reference's legacy typer
//   (etc/hisatgenotype_hla_cyp.py:236-302 joint_abundance), on the GPU.
//
// A pair q = (a, b) is compatible with class k once for every member x of the class whose NAME occurs inside the pair's
// key string "a-b" (the legacy code tests `allele in allele_pair`, a substring test, :280 / :284): the host hands over one
// bitset U_q per pair (alleles whose name is a substring of a's or b's name), so the weight is
//     w[k][q] = popcount(S_k & U_q)                      (pair_weight_kernel: the only contraction of the path, 0/1 x 0/1)
// and one iteration of the legacy loop (:288-297) is
//     prob_k = sum_q w[k][q] P[q];   next[q] = P[q] * sum_k n_k w[k][q] / prob_k;   normalise;   diff = sum |P - next|;
//     keep the pairs with next * 2 > max(next) (choose_top_alleles :259-270); normalise the kept ones.
// The contraction is C x Q x A bit operations once per call (C classes ~ 10^2-10^3, Q kept pairs ~ 10^1-10^4 after the first
// pruning, A alleles): microseconds with 64-bit AND + POPC; nothing here is large enough for tensor cores (binary mma would
// only pay if ncu showed this kernel compute-bound, and it is three orders of magnitude below that).  The loop itself runs
// in ONE CTA: per iteration C x Q multiply-adds on small integers, all sums in a fixed order (reproducible run to run).
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace {

constexpr int PAIR_THREADS = 1024;

__global__ void __launch_bounds__(256) pair_weight_kernel(const uint64_t *__restrict__ cls, const uint64_t *__restrict__ un, int C, int Q,
                                                          int wp, uint8_t *__restrict__ w, int *__restrict__ overflow) {
    const int64_t n = (int64_t)C * Q;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / Q), q = (int)(i % Q);
        const uint64_t *a = cls + (size_t)k * wp, *b = un + (size_t)q * wp;
        int s = 0;
        for (int j = 0; j < wp; j++) s += __popcll(a[j] & b[j]);
        if (s > 255) {
            *overflow = 1;
            s = 255;
        }
        w[i] = (uint8_t)s;
    }
}

__device__ __forceinline__ double pair_block_sum(double v, double *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double x = red[lane];
        x = warp_sum(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double pair_block_max(double v, double *red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double x = red[lane];
        x = warp_max(x);
        if (lane == 0) red[32] = x;
    }
    __syncthreads();
    const double r = red[32];
    __syncthreads();
    return r;
}

// One CTA runs the whole loop.  P / next: [Q]; alive: [Q]; coef: [C].
__global__ void __launch_bounds__(PAIR_THREADS, 1)
    pair_em_kernel(const uint8_t *__restrict__ w, const double *__restrict__ cnt, int C, int Q, double *P, double *next,
                   uint8_t *alive, double *coef, int32_t *iters_out) {
    __shared__ double red[33];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double diff = 1.0;
    int iter = 0;
    while (diff > 0.0001 && iter < 1000) {
        // prob_k and coef_k = n_k / prob_k (a class no kept pair touches contributes nothing: the legacy loop never divides)
        for (int k = warp; k < C; k += PAIR_THREADS / 32) {
            const uint8_t *row = w + (size_t)k * Q;
            double s = 0.0;
            for (int q = lane; q < Q; q += 32)
                if (alive[q]) s += (double)row[q] * P[q];
            s = warp_sum(s);
            if (lane == 0) coef[k] = s > 0.0 ? cnt[k] / s : 0.0;
        }
        __syncthreads();
        // next[q]; a pair no class touches is absent from the legacy dict: marked by a negative value
        double part = 0.0;
        for (int q = tid; q < Q; q += PAIR_THREADS) {
            double v = -1.0;
            if (alive[q]) {
                double acc = 0.0;
                bool touched = false;
                for (int k = 0; k < C; k++) {
                    const uint8_t x = w[(size_t)k * Q + q];
                    if (x) {
                        touched = true;
                        acc += coef[k] * (double)x;
                    }
                }
                if (touched) {
                    v = P[q] * acc;
                    part += v;
                }
            }
            next[q] = v;
        }
        const double total = pair_block_sum(part, red);
        double dpart = 0.0, best = 0.0;
        for (int q = tid; q < Q; q += PAIR_THREADS) {
            if (!alive[q]) continue;
            double v = next[q];
            if (v >= 0.0) {
                v /= total;
                next[q] = v;
                dpart += fabs(P[q] - v);
                best = fmax(best, v);
            } else {
                dpart += P[q];
            }
        }
        diff = pair_block_sum(dpart, red);
        best = pair_block_max(best, red);
        // choose_top_alleles: keep next * 2 > best, renormalise
        double kpart = 0.0;
        for (int q = tid; q < Q; q += PAIR_THREADS) {
            const double v = next[q];
            const bool keep = alive[q] && v >= 0.0 && !(v * 2.0 <= best);
            alive[q] = keep ? 1 : 0;
            if (keep) kpart += v;
        }
        const double ktotal = pair_block_sum(kpart, red);
        for (int q = tid; q < Q; q += PAIR_THREADS) P[q] = alive[q] ? next[q] / ktotal : 0.0;
        __syncthreads();
        iter++;
    }
    if (tid == 0) *iters_out = iter;
}

struct Scratch {
    void *p = nullptr;
    ~Scratch() {
        if (p) cudaFree(p);
    }
};

}  // namespace

extern "C" int hgt_pair_em(hgt_ctx *ctx, const uint64_t *class_bits, const int64_t *class_count, int32_t n_classes, int32_t wp,
                           const uint64_t *pair_bits, int32_t n_pairs, const double *p0, double *prob, int32_t *iters) {
    if (!ctx || !class_bits || !class_count || !pair_bits || !p0 || !prob || n_classes < 1 || n_pairs < 1 || wp < 1) {
        hgt_set_error("hgt_pair_em: bad argument");
        return HGT_ERR_ARG;
    }
    HGT_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t C = (size_t)n_classes, Q = (size_t)n_pairs;
    auto a256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_cls = 0, o_un = a256(o_cls + C * wp * 8), o_cnt = a256(o_un + Q * wp * 8), o_w = a256(o_cnt + C * 8),
                 o_p = a256(o_w + C * Q), o_next = a256(o_p + Q * 8), o_alive = a256(o_next + Q * 8),
                 o_coef = a256(o_alive + Q), o_misc = a256(o_coef + C * 8), bytes = o_misc + 256;
    Scratch s;
    HGT_CUDA(cudaMalloc(&s.p, bytes));
    unsigned char *d = static_cast<unsigned char *>(s.p);
    std::vector<double> cnt(C);
    for (size_t k = 0; k < C; k++) cnt[k] = (double)class_count[k];
    std::vector<uint8_t> alive(Q);
    for (size_t q = 0; q < Q; q++) alive[q] = p0[q] > 0.0 ? 1 : 0;
    HGT_CUDA(cudaMemcpyAsync(d + o_cls, class_bits, C * wp * 8, cudaMemcpyHostToDevice, st));
    HGT_CUDA(cudaMemcpyAsync(d + o_un, pair_bits, Q * wp * 8, cudaMemcpyHostToDevice, st));
    HGT_CUDA(cudaMemcpyAsync(d + o_cnt, cnt.data(), C * 8, cudaMemcpyHostToDevice, st));
    HGT_CUDA(cudaMemcpyAsync(d + o_p, p0, Q * 8, cudaMemcpyHostToDevice, st));
    HGT_CUDA(cudaMemcpyAsync(d + o_alive, alive.data(), Q, cudaMemcpyHostToDevice, st));
    HGT_CUDA(cudaMemsetAsync(d + o_misc, 0, 256, st));
    const unsigned grid = (unsigned)std::min<size_t>((C * Q + 255) / 256, (size_t)ctx->sm_count * 8);
    pair_weight_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint64_t *>(d + o_cls), reinterpret_cast<const uint64_t *>(d + o_un),
                                             n_classes, n_pairs, wp, d + o_w, reinterpret_cast<int *>(d + o_misc));
    pair_em_kernel<<<1, PAIR_THREADS, 0, st>>>(d + o_w, reinterpret_cast<const double *>(d + o_cnt), n_classes, n_pairs,
                                               reinterpret_cast<double *>(d + o_p), reinterpret_cast<double *>(d + o_next), d + o_alive,
                                               reinterpret_cast<double *>(d + o_coef), reinterpret_cast<int32_t *>(d + o_misc + 64));
    ctx->launches += 2;
    HGT_CUDA(cudaGetLastError());
    int32_t misc[32];
    HGT_CUDA(cudaMemcpyAsync(prob, d + o_p, Q * 8, cudaMemcpyDeviceToHost, st));
    HGT_CUDA(cudaMemcpyAsync(misc, d + o_misc, sizeof(misc), cudaMemcpyDeviceToHost, st));
    HGT_CUDA(cudaStreamSynchronize(st));
    if (misc[0]) {
        hgt_set_error("hgt_pair_em: a class shares more than 255 alleles with one pair's name set");
        return HGT_ERR_UNSUPPORTED;
    }
    if (iters) *iters = misc[16];
    return HGT_OK;
}
