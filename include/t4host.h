/*
 * t4host.h — C-ABI of libt4host.so: the host-side mirror of the reference's Tensor / Model
 * class surface (src/mu/tensor.h:51-190, src/nn/model.h:36-164) re-implemented on top of the
 * kernel C-ABI (include/t4k.h).  The C++ classes (t4::Tensor, t4::Model — same method names and
 * argument meaning as the reference) live in tensorforth_b200/csrc/host/t4host.hpp; this flat
 * C interface is what Python (ctypes) tests, bench.py and the Forth front-end bind to.
 *
 * Handles are opaque pointers.  All tensors are FP32 NHWC on the current CUDA device, allocated
 * from the CUDA stream-ordered memory pool (replaces MMU::talloc over a 2 GB managed TLSF arena,
 * src/mu/mmu.cu:37-66,199-209; no 2 GB / 32-bit-offset limit).  Calls are asynchronous on one
 * library stream; functions that return a host scalar or copy to host synchronise that stream
 * (the reference synchronises after every kernel, src/ten4_types.h:192).
 * Errors: functions return 0 / a valid handle on success; on failure a negative T4K_E* code (or
 * NULL) and t4h_last_error() holds the reference-style message ("tensor#ten_op ... dim?").
 */
#ifndef T4HOST_H
#define T4HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct t4h_tensor_s *t4h_tensor;
typedef struct t4h_model_s  *t4h_model;

const char *t4h_last_error(void);
int   t4h_init(int device);                       /* cudaSetDevice + stream + mempool; idempotent */
void *t4h_stream(void);                           /* the library's cudaStream_t */
int   t4h_sync(void);
long  t4h_launch_count(void);

/* ---- Tensor (src/mu/tensor.h, words `vector matrix tensor` of src/vm/tenvm.cpp:458-470) ---- */
t4h_tensor t4h_tensor_new(int rank, uint32_t n, uint32_t h, uint32_t w, uint32_t c);   /* rank 1: h=numel */
void  t4h_tensor_free(t4h_tensor t);
float *t4h_tensor_data(t4h_tensor t);             /* device pointer */
int   t4h_tensor_shape(t4h_tensor t, uint32_t nhwc[4]);
int64_t t4h_tensor_numel(t4h_tensor t);
int   t4h_tensor_rank(t4h_tensor t);
int   t4h_tensor_h2d(t4h_tensor t, const float *src, int64_t n);    /* `={` / t! */
int   t4h_tensor_d2h(t4h_tensor t, float *dst, int64_t n);          /* `.` / t@  (syncs) */
int   t4h_tensor_reshape(t4h_tensor t, int rank, uint32_t n, uint32_t h, uint32_t w, uint32_t c);
t4h_tensor t4h_tensor_copy(t4h_tensor t);                           /* MMU::copy, `copy` word */
int   t4h_tensor_map(t4h_tensor t, int op, float v);                /* Tensor::map, words exp ln relu fill ... */
int   t4h_tensor_identity(t4h_tensor t);                            /* `eye` */
int   t4h_tensor_rand(t4h_tensor t, int opt);                       /* `rand` / `randn` */
int   t4h_ten_op_s(int op, t4h_tensor A, float v, t4h_tensor O);    /* Tensor::ten_op(op,A,v,O) */
int   t4h_ten_op_t(int op, t4h_tensor A, t4h_tensor B, t4h_tensor O); /* Tensor::ten_op(op,A,B,O) incl. N-broadcast */
int   t4h_mm(t4h_tensor A, t4h_tensor B, t4h_tensor O, int inc, int tA, int tB);      /* Tensor::mm → gemm3 */
int   t4h_gemm(int variant, t4h_tensor A, t4h_tensor B, t4h_tensor O, float alpha, float beta, int tA, int tB); /* gemm1..4 */
t4h_tensor t4h_matmul(t4h_tensor A, t4h_tensor B);                  /* `@`: TensorVM::_tdot rank rules; NULL on "dim?" */
t4h_tensor t4h_transpose(t4h_tensor A);                             /* `transpose` */
float t4h_tensor_sum(t4h_tensor t);
float t4h_tensor_avg(t4h_tensor t);
float t4h_tensor_std(t4h_tensor t);
float t4h_tensor_norm(t4h_tensor t);
float t4h_tensor_max(t4h_tensor t);
float t4h_tensor_min(t4h_tensor t);
float t4h_tensor_dot(t4h_tensor A, t4h_tensor B);
float t4h_tensor_loss(t4h_tensor out_copy, int loss_op, t4h_tensor tgt);   /* Tensor::loss (non-destructive here) */

/* ---- Model (src/nn/model.h; words of src/vm/netvm.cpp:291-485) ---- */
t4h_model t4h_model_new(uint32_t n, uint32_t h, uint32_t w, uint32_t c);   /* `nn.model` */
void  t4h_model_free(t4h_model m);
/* Model::add(fn, n, bias, opt): opt = {kernel, stride, padding, dilation} for conv (may be NULL) */
int   t4h_model_add(t4h_model m, int layer, uint32_t n, float bias, const uint16_t *opt);
int   t4h_model_numel(t4h_model m);                                  /* layer tensors incl. output */
t4h_tensor t4h_model_layer(t4h_model m, int i);                      /* `n@` (negative index allowed); borrowed */
t4h_tensor t4h_model_param(t4h_model m, int i, int which);           /* `nn.w nn.b nn.dw nn.db nn.ex` which=0..4; borrowed */
int   t4h_model_set_param(t4h_model m, int i, int which, t4h_tensor t);   /* `nn.w=` `nn.b=` (copies t) */
int   t4h_model_train(t4h_model m, int on);                          /* `trainable` */
int   t4h_model_fuse(t4h_model m, int on);                           /* multi-layer fused kernels on (default) / off: same tensors either way */
int   t4h_model_forward(t4h_model m, t4h_tensor input);              /* `forward` */
int   t4h_model_backprop(t4h_model m, t4h_tensor tgt);               /* `backprop` (tgt NULL → cached one-hot) */
float t4h_model_loss(t4h_model m, int loss_op, t4h_tensor tgt);      /* `loss.mse|bce|ce|nll` (syncs) */
int   t4h_model_loss_async(t4h_model m, int loss_op, t4h_tensor tgt, float *loss_dev);
/* ---- Dataset (src/mu/dataset.h, dataset.cu): mini-batch feeding.  The loader side (file parsing, src/ld) stays with the
 * caller; what it would hand to Dataset::_load — the raw U8 image block and U8 labels of a mini-batch — goes to
 * t4h_dataset_stage (async H2D of the BYTES on a copy stream, double buffered: stage batch i+1 while batch i trains);
 * t4h_dataset_commit normalises on the device into the dataset's tensor ((u8 - mean) * (1/scale), dataset.cu:142). */
typedef void *t4h_dataset;
t4h_dataset t4h_dataset_create(int n, int h, int w, int c);
void  t4h_dataset_destroy(t4h_dataset d);
void  t4h_dataset_normalize(t4h_dataset d, float mean, float scale);          /* word `normalize` ( DS mean scale -- DS' ) */
int   t4h_dataset_stage(t4h_dataset d, const uint8_t *img_host, const uint8_t *lab_host, int n);
int   t4h_dataset_commit(t4h_dataset d);
t4h_tensor t4h_dataset_tensor(t4h_dataset d);                                 /* the dataset as the Tensor it is */
const int32_t *t4h_dataset_labels(t4h_dataset d);                             /* device int32 labels of the committed batch */
int   t4h_model_forward_ds(t4h_model m, t4h_dataset d);                       /* Model::forward(Dataset&): + onehot + hit (forward.cu:72-75) */
/* one iteration of `ds for forward loss backprop nn.adam next`: commit + one-hot + the captured train step */
int   t4h_model_step_graph_ds(t4h_model m, t4h_dataset d, int loss_op, float *loss_dev, int optimizer, float lr, float b1, float b2, float wd);
/* the same with the loss read-back pipelined by one step: this step's loss goes to pinned host memory asynchronously,
 * *prev_loss receives the PREVIOUS call's loss (NaN on the first call); t4h_model_train_flush waits for the last one.
 * One host call per training iteration: nothing else touches the GPU queue. */
int   t4h_model_train_step_ds(t4h_model m, t4h_dataset d, int loss_op, float *loss_dev, int optimizer, float lr, float b1, float b2, float wd, float *prev_loss);
int   t4h_model_train_flush(t4h_model m, float *last_loss);
int   t4h_model_onehot_labels(t4h_model m, const int32_t *labels_dev);   /* Model::onehot(Dataset&) on device */
int   t4h_model_onehot_set(t4h_model m, t4h_tensor hot);             /* `nn.onehot=` */
int   t4h_model_hit(t4h_model m, int recalc);                        /* `nn.hit` */
int   t4h_model_sgd(t4h_model m, float lr, float b);                 /* `nn.sgd` */
int   t4h_model_adam(t4h_model m, float lr, float b1, float b2);     /* `nn.adam` */
int   t4h_model_adamw(t4h_model m, float lr, float wd, float b1, float b2);
/* flat parameter arenas (built at the first optimizer call): pointers + float count; DG is what a
 * data-parallel caller sum-allreduces between backprop and the optimizer */
/* generic CUDA-graph capture of a sequence of library calls (e.g. one GAN iteration over two models): begin, make the calls (no host
 * reads: they synchronise), end -> handle; replay with t4h_graph_launch.  Warm the sequence up once before capturing. */
int   t4h_capture_begin(void);
void *t4h_capture_end(void);
int   t4h_graph_launch(void *graph);
void  t4h_graph_free(void *graph);
/* words `save` / `load` on a model (src/vm/netvm.cpp:479-480 -> src/io/aio_model.cpp): the reference's model file — text header and
 * layer lines, then `--- w.<layer>` / `--- b.<layer>` sections of raw FP32.  load fills an already built model (parameter path). */
int   t4h_model_save(t4h_model m, const char *fname);
int   t4h_model_load(t4h_model m, const char *fname);      /* parameters; + optimizer state when the file carries it */
int   t4h_model_save_state(t4h_model m, const char *fname); /* t4h_model_save + the optimizer state (moment arenas, step count) for resume, appended behind
                                                             * the reference's closing section: the reference's reader still loads the file */
int   t4h_model_arena(t4h_model m, float **G, float **DG, int64_t *total);
/* capture forward+loss+backprop+optimizer into one CUDA graph and replay it (launch-bound regime);
 * optimizer: 0 sgd, 1 sgd+momentum, 2 adam, 3 adamw, -1 none (data parallel: all-reduce DG, then call the optimizer) */
/* data parallel: attach a connected t4k_comm_t (include/t4k.h); from then on sgd/adam/adamw — also inside step_graph —
 * sum the gradient arena over the ranks inside the optimizer kernel; scal[0..nscal) device floats ride along (summed) */
int   t4h_model_dp_attach(t4h_model m, void *comm, float *scal, int nscal);
/* this model holds shard `rank` of `world` equal shards of the global batch: dropout masks come from the shard's global element offsets, batch-norm
 * statistics are summed over the ranks on `comm_stat` (a t4k_comm_t of its own, capacity >= 4 x t4h_model_bn_channels; may be NULL without batchnorm).
 * A model with batchnorm layers refuses t4h_model_dp_attach until this was called with a communicator. */
int   t4h_model_dp_shard(t4h_model m, int rank, int world, void *comm_stat);
int   t4h_model_bn_channels(t4h_model m);        /* widest batchnorm layer (0: none) */
int   t4h_tensor_rand_sharded(t4h_tensor t, int opt, int rank, int world);   /* t4k_rand_sharded on a batch-major tensor holding shard `rank` */
void *t4h_side_stream(void);                     /* the side stream of the current lane (work forked inside a step) */
int   t4h_use_lane(int lane);
int   t4h_set_dp_rest(int on);                   /* 1 (default): the exchange + optimizer of everything past the first chunk runs on the side stream under the
                                                  * first layer's finish launch; 0: one exchange launch at the end of the step (also T4K_DP_REST=0) */
int   t4h_set_dp_early(int mode);                /* early half of the data-parallel exchange inside the captured step: 0 push kernel, 1 copy engines + early
                                                  * exchange of the rest of the arena, 2 whole early exchange, 3 by world size (default; also T4K_DP_EARLY=sm|dma|range) */                    /* tests: switch the process to stream set `lane` (0..3); lane 0 is the default */
int   t4h_model_step_graph(t4h_model m, t4h_tensor input, t4h_tensor tgt, int loss_op, float *loss_dev,
                           int optimizer, float lr, float b1, float b2, float wd);

#ifdef __cplusplus
}
#endif
#endif /* T4HOST_H */
