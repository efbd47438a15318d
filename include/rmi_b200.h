/* rmi_b200.h — C ABI of the B200-native two-layer RMI trainer (librmi_b200.so).
 *
 * Drop-in boundary for the reference's `rmi_lib::train`
 *   pub fn train<T: TrainingKey>(data: &RMITrainingData<T>, model_spec: &str,
 *                                branch_factor: u64) -> TrainedRMI
 * (reference rmi_lib/src/train/mod.rs:100-126, re-exported rmi_lib/src/lib.rs:10; callers
 * src/main.rs:203,276, rmi_lib/src/optimizer.rs:227, train/mod.rs:146,174).  The reference
 * has no FFI layer of its own; a Rust `rmi_lib` would bind these symbols with `extern "C"`
 * in place of `two_layer::train_two_layer` (train/two_layer.rs:101) — INTEGRATION.md shows
 * the stub.  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Conventions (mirroring the reference's):
 *   - input  : an immutable, shareable key set (RMITrainingData, models/mod.rs:233-317).
 *              Here: `rmi_dataset`, device-resident, read-only, may be used by concurrent
 *              rmi_train calls (the optimizer does that, optimizer.rs:224).
 *   - output : an owned TrainedRMI (train/mod.rs:18-33).  Here: `rmi_result`, host memory
 *              owned by the library, released with rmi_result_free.
 *   - errors : the reference panics (process abort).  Here every entry point returns an
 *              rmi_status; rmi_last_error() gives the message the reference would have
 *              printed.  The library never calls exit/abort.
 * Keys must be sorted ascending (the reference's file format requires it, README.md:26-31);
 * an unsorted array is reported as RMI_ERR_PANIC ("keys are not sorted").
 */
#ifndef RMI_B200_H_
#define RMI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/load.rs:15-19 DataType / rmi_lib/src/models/mod.rs:41-43 KeyType */
typedef enum { RMI_KEY_U64 = 0, RMI_KEY_U32 = 1, RMI_KEY_F64 = 2 } rmi_key_type;

typedef enum {
  RMI_OK = 0,
  RMI_ERR_PANIC = 1,        /* the reference would have panicked (assert!/unwrap/panic!) */
  RMI_ERR_INVALID = 2,      /* bad argument at this boundary (null pointer, bad enum, ...) */
  RMI_ERR_CUDA = 3,         /* CUDA runtime / driver failure, or no usable device */
  RMI_ERR_UNSUPPORTED = 4   /* valid in the reference, not offered by this build */
} rmi_status;

/* Model identifiers (reference train/mod.rs:37-54 train_model name table).  RADIX_TABLE
 * covers radix8/18/22/26/28 (table bits in rmi_result.l0_table_bits); BRADIX reports its
 * clamp-high / clamp-low variant in rmi_result.l0_bradix_high. */
typedef enum {
  RMI_MODEL_LINEAR = 0, RMI_MODEL_ROBUST_LINEAR = 1, RMI_MODEL_LINEAR_SPLINE = 2, RMI_MODEL_CUBIC = 3,
  RMI_MODEL_LOGLINEAR = 4, RMI_MODEL_NORMAL = 5, RMI_MODEL_LOGNORMAL = 6, RMI_MODEL_RADIX = 7,
  RMI_MODEL_RADIX_TABLE = 8, RMI_MODEL_BRADIX = 9, RMI_MODEL_HISTOGRAM = 10
} rmi_model_id;

/* rmi_train flags */
enum {
  RMI_FLAG_STATS_ONLY = 1u,     /* do not copy leaf parameters/errors to the host (optimizer use:
                                   only the statistics are consumed, optimizer.rs:163-171) */
  RMI_FLAG_TOP_FIT_EXACT = 2u,  /* fit linear/robust_linear/loglinear/normal/lognormal TOP models
                                   with the reference's order-dependent serial recurrence
                                   (linear.rs:12-59) on one device thread: bit-identical to the
                                   reference, seconds at 200 M keys.  Default is the parallel fit
                                   (tree reduction, coefficients equal within 1e-9 relative). */
  RMI_FLAG_NO_ERRORS = 4u       /* reserved for --no-errors (main.rs:84-86); errors are still computed */
};

/* A device-resident sorted key set.  Replaces src/load.rs:132-157 load_data + the mmap
 * slice adapters (load.rs:21-95): same content (n packed little-endian keys), in HBM. */
typedef struct rmi_dataset rmi_dataset;

/* Copy n host keys to `device` (cudaMemcpyAsync from the caller's buffer; pinned buffers
 * transfer at full PCIe rate).  host_keys must stay valid until the call returns. */
int rmi_dataset_create(const void* host_keys, uint64_t n, rmi_key_type key_type, int device,
                       rmi_dataset** out);
/* Borrow keys that already live in device memory on `device` (no copy, caller keeps ownership
 * and must keep them alive and unmodified while the dataset exists). */
int rmi_dataset_wrap_device(const void* device_keys, uint64_t n, rmi_key_type key_type, int device,
                            rmi_dataset** out);
/* Read a reference-format key file (u64 LE count + packed keys, README.md:26-31; key type from
 * the file-name suffix as src/main.rs:122-132 does when key_type < 0) straight into HBM through
 * a double-buffered pinned staging ring. */
int rmi_dataset_load_file(const char* path, int key_type_or_negative, int device, rmi_dataset** out);
uint64_t rmi_dataset_len(const rmi_dataset* ds);
int rmi_dataset_key_type(const rmi_dataset* ds);
void rmi_dataset_destroy(rmi_dataset* ds);

/* Mirror of TrainedRMI (reference train/mod.rs:18-33).  All pointers are host memory owned
 * by the result; they stay valid until rmi_result_free. */
typedef struct {
  uint64_t num_rmi_rows;        /* TrainedRMI.num_rmi_rows  */
  uint64_t num_data_rows;       /* TrainedRMI.num_data_rows */
  uint64_t branching_factor;    /* TrainedRMI.branching_factor */
  double model_avg_error;       /* two_layer.rs:274-275 */
  double model_avg_l2_error;    /* two_layer.rs:277-279 */
  double model_avg_log2_error;  /* two_layer.rs:281-282 */
  double model_max_log2_error;  /* two_layer.rs:284 */
  uint64_t model_max_error;     /* two_layer.rs:267-271 */
  uint64_t model_max_error_idx;
  uint64_t build_time_ns;       /* wall clock of the rmi_train call (train/mod.rs:103,114-118) */
  uint64_t device_time_ns;      /* CUDA-event time of the kernels of this build */
  uint64_t phase_device_ns[4];  /* CUDA-event time per phase: [0] top-model fit, [1] leaf boundaries,
                                   [2] fused leaf fit + forward/error pass, [3] statistics */

  /* layer 0: TrainedRMI.rmi[0][0] */
  uint32_t l0_model_id;         /* rmi_model_id */
  uint32_t l0_bradix_high;      /* bradix: 1 = bradix_clamp_high, 0 = bradix_clamp_low */
  uint32_t l0_table_bits;       /* radix table: 8/18/22/26/28 */
  uint32_t l0_num_fparams;      /* float parameters in Model::params() order */
  double l0_fparams[4];
  uint32_t l0_num_iparams;      /* integer parameters in Model::params() order */
  uint32_t _pad0;
  uint64_t l0_iparams[4];
  uint64_t l0_table32_len;      /* radix table: hint table (ModelParam::Int32Array) */
  const uint32_t* l0_table32;
  uint64_t l0_array1_len;       /* histogram: radix index (ModelParam::IntArray) */
  const uint64_t* l0_array1;
  uint64_t l0_array2_len;       /* histogram: pivots (ModelParam::IntArray) */
  const uint64_t* l0_array2;

  /* layer 1: TrainedRMI.rmi[1][0..N] and TrainedRMI.last_layer_max_l1s */
  uint32_t l1_model_id;
  uint32_t l1_params_per_model;
  const double* l1_params;      /* N x params_per_model, leaf order; NULL with STATS_ONLY */
  const uint64_t* l1_errors;    /* N; NULL with STATS_ONLY */
  const uint64_t* l1_counts;    /* N keys-per-leaf as counted by two_layer.rs:207-217; NULL with STATS_ONLY */
  uint32_t could_not_replace;   /* two_layer.rs:199-202 warning condition */
  uint32_t top_fit_exact;       /* 1 if the top model came from the serial recurrence */
} rmi_result;

/* rmi_lib::train.  model_spec is "top,leaf" (train/mod.rs:104-109); only two-layer specs are
 * accepted, as in the reference (train/mod.rs:123-125). */
int rmi_train(const rmi_dataset* ds, const char* model_spec, uint64_t branch_factor, uint32_t flags,
              rmi_result** out);
/* As rmi_train, but the top model's float parameters are given instead of fitted (linear,
 * robust_linear, linear_spline: alpha,beta; cubic: a,b,c,d; normal/lognormal: mean,stdev,scale). */
int rmi_train_with_top(const rmi_dataset* ds, const char* model_spec, uint64_t branch_factor, uint32_t flags,
                       const double* l0_fparams, uint32_t n_fparams, rmi_result** out);
void rmi_result_free(rmi_result* r);

/* Message of the last failure on the calling thread ("" if none). */
const char* rmi_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rmi_kernel_launch_count(void);
/* Library / build identification, e.g. "rmi_b200 0.1 sm_100a". */
const char* rmi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RMI_B200_H_ */
