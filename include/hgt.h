/* libhgt — C ABI of the B200-native HISAT-genotype typing hot path.
 *
 * The reference (DaehwanKimLab/hisat-genotype v1.3.2) is pure Python and has no FFI for this path; the
 * seam is the module-level Python API.  Each entry point below names the reference code it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only; caller-owned buffers; every function returns 0 on success or a
 * negative hgt_status; hgt_last_error() gives a message for the calling thread.  One hgt_ctx per (process,
 * device); a context is not thread-safe, distinct contexts are independent.  Functions whose name ends in
 * `_dev` take DEVICE pointers and a CUDA stream and never synchronise; all others take HOST pointers, do their
 * own host<->device copies and return when the result is in the caller's buffers.
 *
 * Bit layout: allele sets are arrays of uint64 words, bit (a & 63) of word (a >> 6) = allele a, where alleles
 * of a locus are numbered in sorted-name order with the backbone excluded (the order of names inside a
 * Gene_cmpt key, reference hisatgenotype_typing_core.py:1229-1230).  Row pitch `wp` is in 64-bit words and
 * must be even (16-byte rows for bulk copies); use hgt_row_pitch(n_alleles).
 */
#ifndef HGT_H
#define HGT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hgt_ctx hgt_ctx;
typedef struct hgt_locus hgt_locus;
typedef struct hgt_typing hgt_typing;

typedef enum {
    HGT_OK = 0,
    HGT_ERR_CUDA = -1,          /* CUDA runtime error (message has the call site) */
    HGT_ERR_KEY = -2,           /* reference would raise KeyError in the SQUAREM step (common:1365-1369) */
    HGT_ERR_ZERODIV = -3,       /* reference would raise ZeroDivisionError in normalize (common:1285-1297) */
    HGT_ERR_ARG = -4,           /* bad argument */
    HGT_ERR_UNSUPPORTED = -5,   /* shape outside what the kernels are built for (message says which) */
    HGT_ERR_PARSE = -6,         /* malformed alignment record (the reference would assert) */
    HGT_ERR_AMBIGUITY = -7,     /* check_amb_uniqueness would exit(1) (validation_check.py:313-341) */
    HGT_ERR_NOMEM = -8
} hgt_status;

const char *hgt_last_error(void);
int hgt_abi_version(void);
/* words per allele-set row (even, >= 2) for n_alleles */
int hgt_row_pitch(int n_alleles);

/* ---- context ------------------------------------------------------------------------------------------- */
int hgt_init(int device, hgt_ctx **out);
void hgt_free(hgt_ctx *ctx);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
int64_t hgt_launch_count(const hgt_ctx *ctx);
int hgt_sm_count(const hgt_ctx *ctx);

/* ---- stage (b): EM abundance ----------------------------------------------------------------------------
 * Replaces single_abundance(Gene_cmpt, remove_low_abundance_allele, Gene_length)
 *   reference hisatgenotype_modules/hisatgenotype_typing_common.py:1282-1410 (+ prob_diff :1272-1279).
 * class_bits [n_classes][wp]: member alleles of each Gene_cmpt key, classes in dict (first-seen) order;
 * class_count[n_classes]: the dict values; allele_len[n_alleles] or NULL (Gene_length == {}).
 * Outputs: prob[n_alleles] (0 for alleles not in the result), in_result[n_alleles] (1 if the allele is a key
 * of the reference's final Gene_prob), first_class[n_alleles] (index of the first class that inserted the
 * allele into the final dict — the tie-break of the reference's stable sort), iters (loop count, :1351). */
int hgt_em(hgt_ctx *ctx, const uint64_t *class_bits, const int64_t *class_count, int32_t n_classes,
           int32_t n_alleles, int32_t wp, const double *allele_len, int32_t remove_low, double *prob,
           uint8_t *in_result, int32_t *first_class, int32_t *iters);

/* Batched form: n_problems independent EM problems in one launch (one CTA per problem).  Problem i uses
 * classes [class_off[i], class_off[i+1]) of the concatenated class arrays and alleles
 * [allele_off[i], allele_off[i+1]) of the concatenated per-allele arrays; all rows share pitch wp.
 * status[i] receives the per-problem hgt_status. */
int hgt_em_batch(hgt_ctx *ctx, int32_t n_problems, const uint64_t *class_bits, const int64_t *class_count,
                 const int64_t *class_off, const int64_t *allele_off, int32_t wp, const double *allele_len,
                 const uint8_t *remove_low, double *prob, uint8_t *in_result, int32_t *first_class,
                 int32_t *iters, int32_t *status);

/* Device-pointer form of hgt_em for inputs already resident in HBM (class_count as double).  `workspace` must
 * hold hgt_em_workspace_bytes() bytes.  n_ctas = 0 lets the library choose (1 CTA for small problems, one per
 * SM for large ones).  fixed_iters > 0 runs exactly that many loop iterations ignoring convergence (bench). */
size_t hgt_em_workspace_bytes(const hgt_ctx *ctx, int32_t n_classes, int32_t n_alleles);
int hgt_em_dev(hgt_ctx *ctx, void *stream, const uint64_t *class_bits, const double *class_count,
               int32_t n_classes, int32_t n_alleles, int32_t wp, const double *allele_len, int32_t remove_low,
               int32_t fixed_iters, int32_t n_ctas, double *prob, uint8_t *in_result, int32_t *first_class,
               int32_t *iters_status /* [2]: iters, status */, void *workspace);

#ifdef __cplusplus
}
#endif
#endif /* HGT_H */
