// t4host.hpp — host-side mirror of the reference's Tensor / Model class surface
//   t4::Tensor  ←→  src/mu/tensor.h:51-190   (struct Tensor : T4Base)
//   t4::Model   ←→  src/nn/model.h:36-164    (class Model : T4Base)
// Same method names, argument order and meaning; the bodies call the kernel C-ABI (include/t4k.h)
// instead of launching FORK*() kernels + cudaDeviceSynchronize().  Everything above this file in the
// reference (Forth VMs, MMU/TLSF, IO, TensorBoard) is a caller of exactly these methods.
#pragma once
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../../../include/t4k.h"

namespace t4 {

typedef float    DU;
typedef uint64_t U64;
typedef uint32_t U32;
typedef uint16_t U16;
typedef int32_t  S32;

typedef enum t4k_math_op math_op;
typedef enum t4k_layer   t4_layer;
typedef enum t4k_loss    t4_loss;      // `enum` tag: t4k_loss is also the C-ABI function name
typedef enum { OPTI_SGD = 0, OPTI_SGDM, OPTI_ADAM, OPTI_ADAMW } t4_optimizer;   // src/nn/ntypes.h:57-62

#define T4_DIM_SZ 16
#define DU_EPS_H  1.0e-6f

// process-wide runtime: device, stream, error slot (src/ten4.cu:125-176 keeps these in TensorForth)
struct Runtime {
    static int   init(int device);
    static void *stream();
    static int   sync();
    static void  error(const char *fmt, ...);          // ERROR(...) of src/ten4_types.h:25
    static const char *last_error();
    static void *alloc(size_t bytes);                  // stream-ordered pool (replaces MMU::talloc)
    static void  free(void *p);
    static int   use_lane(int k);                      // switch the process to stream set k (tests: several ranks of one process)
};

struct Tensor {
    // ---- T4Base (src/t4base.h:50-115)
    U64      numel = 0;
    U32      rank  = 0;
    U32      iparm = 0;
    bool     train = false, err = false;
    DU       xparm = 0;
    DU      *data  = nullptr;          ///< device memory (numel + 1 floats: _tmp scratch at data[numel])
    // ---- Tensor (src/mu/tensor.h:52-58)
    U16      stride[4] = {1, 1, 1, 1}; ///< [strideH, strideW, paddingH, paddingW]
    U32      shape[4]  = {1, 1, 1, 1}; ///< HWCN
    t4_layer grad_fn   = T4K_L_NONE;
    Tensor  *grad[5]   = {nullptr, nullptr, nullptr, nullptr, nullptr};
    Tensor  *mtum[5]   = {nullptr, nullptr, nullptr, nullptr, nullptr};
    DU      *_tmp      = nullptr;
    bool     owns      = true;         ///< false for views into a parameter arena

    // life cycle (MMU::tensor / MMU::free, src/mu/mmu.cu:211-268)
    static Tensor &create(U64 sz);
    static Tensor &create(U32 h, U32 w);
    static Tensor &create(U32 n, U32 h, U32 w, U32 c);
    static Tensor &create_like(Tensor &t);
    static Tensor &copy_of(Tensor &t);                 ///< MMU::copy (hard copy)
    static void    destroy(Tensor &t);

    // static ops — result tensor is the last parameter and is returned (tensor.h:59-80)
    static Tensor &ten_op(math_op op, Tensor &A, DU v, Tensor &O);
    static Tensor &ten_op(math_op op, Tensor &A, Tensor &B, Tensor &O);
    static Tensor &dot(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta);
    static Tensor &mm(Tensor &A, Tensor &B, Tensor &O, bool inc = 0, bool tA = 0, bool tB = 0);
    static Tensor &linear(Tensor &A, Tensor &B, Tensor &O, int H, int W, int K, DU alpha, DU beta, bool tA = 0, bool tB = 0);
    static Tensor &gemm1(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA = 0, bool tB = 0);
    static Tensor &gemm2(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA = 0, bool tB = 0);
    static Tensor &gemm3(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA = 0, bool tB = 0);
    static Tensor &gemm4(Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA = 0, bool tB = 0);
    static Tensor &copy(Tensor &A, Tensor &O);
    static Tensor &transpose(Tensor &A, Tensor &T);

    // attributes (tensor.h:106-122)
    U32 &N() { return shape[3]; }
    U32 &H() { return shape[0]; }
    U32 &W() { return shape[1]; }
    U32 &C() { return shape[2]; }
    U64  HWC() { return (U64)shape[0] * shape[1] * shape[2]; }
    U64  size() { return HWC() * N(); }
    DU  *slice(int n) { return &data[HWC() * n]; }
    bool is_same_shape(Tensor &t);

    // arithmetics (tensor.h:126-135); host scalars → these synchronise
    DU  sum();  DU avg();  DU std();  DU norm();  DU max();  DU min();
    DU  dot(Tensor &B);
    DU  loss(t4_loss op, Tensor &tgt);
    U32 has_nan();

    // life-cycle ops (tensor.h:144-153)
    Tensor &reset(void *mem, U64 sz, t4_layer fn = T4K_L_NONE);
    Tensor &reshape(U64 sz);
    Tensor &reshape(U32 h, U32 w);
    Tensor &reshape(U32 n, U32 h, U32 w, U32 c);
    Tensor &identity();
    Tensor &zeros();
    Tensor &map(math_op op, DU v = 0.0f);
    Tensor &normalize(DU avg, DU std);

    // operators (tensor.h:166-180)
    Tensor &operator=(DU v)       { return map(T4K_FILL, v); }
    Tensor &operator+=(DU v)      { return map(T4K_ADD, v); }
    Tensor &operator-=(DU v)      { return map(T4K_SUB, v); }
    Tensor &operator*=(DU v)      { return map(T4K_MUL, v); }
    Tensor &operator=(Tensor &t)  { copy(t, *this); return *this; }
    Tensor &operator+=(Tensor &t) { return ten_op(T4K_ADD, *this, t, *this); }
    Tensor &operator-=(Tensor &t) { return ten_op(T4K_SUB, *this, t, *this); }
    Tensor &operator*=(Tensor &t) { return ten_op(T4K_MUL, *this, t, *this); }

    // host <-> device (replaces the VM's direct pokes into managed memory, tenvm.cpp:535-541)
    int h2d(const DU *h, U64 n = 0);
    int d2h(DU *h, U64 n = 0);
private:
    Tensor &_gemm(int engine, Tensor &A, Tensor &B, Tensor &O, DU alpha, DU beta, bool tA, bool tB, const char *nm);
};

// Dataset (src/mu/dataset.h:14-45): a Tensor whose data is refilled mini-batch by mini-batch, plus the batch's labels.
// The reference converts U8 -> float on the host and copies floats (dataset.cu:124-152); here the U8 block crosses PCIe
// (async, double-buffered staging, own copy stream) and t4k_dataset_load normalises on the device.  The Corpus/loader side
// (file parsing, src/ld) stays with the caller: stage() takes the raw U8 image and label blocks it would hand to _load.
struct Dataset : public Tensor {
    int       batch_sz = 0, batch_id = 0;      ///< dataset.h:17-21
    int32_t  *label = nullptr;                 ///< device int32 labels of the committed batch
    DU        _mean = 0.0f, _scale = 1.0f / 256.0f;   ///< dataset.h:35-36
    static Dataset &create(U32 n, U32 h, U32 w, U32 c);
    static void     destroy(Dataset &d);
    void normalize(DU mean, DU scale);         ///< dataset.cu:33-41
    int  stage(const uint8_t *img_host, const uint8_t *lab_host, int n);   ///< async H2D of the next batch's U8 blocks
    int  commit(DU *hot = nullptr, int E = 0); ///< library stream: wait for the oldest staged batch, normalise into data / label (+ one-hot rows [n,E])
    // commit in three parts, for a caller that folds the normalise launch into its own CUDA graph (Model::step_graph):
    int  commit_begin(const uint8_t **simg, const uint8_t **slab, int *n);     ///< stream-wait for the staged bytes; which staging buffers
    int  commit_launch(const uint8_t *simg, const uint8_t *slab, int n, DU *hot, int E);   ///< the launch itself (capturable)
    void commit_end();                                                          ///< mark the staging buffer consumed
    friend class Model;
private:
    uint8_t *_simg[2] = {nullptr, nullptr}, *_slab[2] = {nullptr, nullptr};
    void    *_staged[2] = {nullptr, nullptr}, *_consumed[2] = {nullptr, nullptr};
    int      _sn[2] = {0, 0};
    unsigned _head = 0, _tail = 0;             ///< staged batches: [_tail, _head)
};

class Model {
    int     _hit  = 0; bool _hit_dev = false;     ///< _hit_dev: the count of the last forward(Dataset&) is still on the device
    int     _iter = 0;
    Tensor *_hot  = nullptr;           ///< cached one-hot vector
    bool    _own_hot = false;
    int    *_cnt_dev = nullptr;        ///< device int for hit()
    // flat parameter arenas (B200-first: one optimizer launch / one all-reduce for the whole model)
    DU     *_G = nullptr, *_DG = nullptr, *_M = nullptr, *_V = nullptr;
    U64     _total = 0;
    void   *_seg_dev = nullptr; int _nseg = 0;
    t4_optimizer _arena_opt = OPTI_SGD;
    // captured train step
    struct StepGraph { void *exec = nullptr; U64 key[13] = {0}; U64 used = 0; } _graphs[6];   // small LRU cache (dataset feeding alternates two staging buffers)
    U64     _graph_clock = 0;
    struct StepExtra { const uint8_t *simg = nullptr, *slab = nullptr; int n = 0; Dataset *ds = nullptr; DU *loss_pin = nullptr; };   // work folded into the captured step
    int     _step_graph(Tensor &input, Tensor &tgt, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd, const StepExtra &x);
    void    _drop_graphs();
    void   *_comm = nullptr; DU *_dp_scal = nullptr; int _dp_nscal = 0;   // data parallel: t4k_comm_t + scalars riding in the exchange
    void   *_comm_stat = nullptr; int _dp_rank = 0, _dp_world = 1;        // data parallel: batch-norm statistics communicator, this rank's shard of the global batch
    int     _second_layer = 0; int64_t _first_end = 0;                    // arena layout: end of the first parameter layer's segments, index of the next parameter layer
    bool    _dp_early = false, _dp_join = false; int64_t _dp_pushed_from = -1;   // split exchange inside step_graph (early push on the side stream)
    void    _dp_push();
    uint32_t _dp_step = 0;             // exchanges issued on _comm so far (= the epoch of every chunk of the arena: selects the slot parity of a DMA push)
    struct DpOptEarly { bool on = false, rest = false, rs = false; int kind = 0; DU lr = 0, b1 = 0, b2 = 0, wd = 0; } _dpo;   // data parallel, inside step_graph: optimizer arguments for the early exchange
    DU     *_pdup = nullptr; bool _want_pdup = false, _pdup_valid = false;   // step_graph: duplicate of the softmax output for the side-stream loss
    const StepExtra *_feed = nullptr; DU *_feed_hot = nullptr;   // step_graph: staged U8 batch still to be loaded (folded into the first fused block when there is one)
    void    _feed_fallback();
    // classifier-head backward deferred onto the side stream (in front of the hidden linear's dW GEMM) while that layer's dX GEMM generates
    // its operand from the forward tensors (t4k_linear_dx_from_head): the arguments of the pending t4k_mlp_head_bwd
    // the TRAIN TAIL inside step_graph (t4k_linear_act_head_train): the forward tail kernel already did the head's backward on its rows; the
    // head's parameter gradients wait as per-CTA partials in _hscratch for t4k_head_grad_finish (side stream)
    Tensor *_fwd_tgt = nullptr; bool _tail_done = false; DU *_hscratch = nullptr; int _hncta = 0;
    struct HeadPending { bool on = false; DU *P, *T, *Ylin, *X2, *F1, *Y1, *W2, *dW2, *dB2, *dB1; int N, E0, E1; } _hp;
    // optimizer split inside step_graph (single GPU): once every gradient but the first parameter layer's is final, the optimizer of the rest of
    // the arena runs on the side stream under the first layers' backward (_opt_push); the fused conv block applies the step to its own filter /
    // bias in the launch that finishes them (t4k_conv_pool_relu_bwd_opt), so the critical path ends one launch earlier
    struct OptEarly { bool on = false, rest = false, first = false, late = false; int kind = 0; DU lr = 0, b1 = 0, b2 = 0, wd = 0; } _oe;
    void    _opt_push();
    bool    _side_join = false, _skip_flat_copy = false;   // backprop: work pending on the side stream / flatten copy already issued there
    DU     *_loss_pin = nullptr; void *_loss_ev[2] = {nullptr, nullptr}; unsigned _tstep = 0;   // train_step read-back ring
    std::vector<Tensor*> _layers;      ///< layer i holds that layer's INPUT; last = output
public:
    int  epoch    = 0;
    DU   max_norm = 0;
    bool train    = true;
    bool err      = false;
    bool fuse     = true;              ///< allow multi-layer fused kernels (same tensors written; off = strict per-layer launches)
    U64  numel() { return _layers.size(); }

    Model(U32 n, U32 h, U32 w, U32 c);                       ///< `nn.model` (netvm.cpp:301-311)
    ~Model();
    void tick() { epoch++; _iter = 0; }
    Tensor &operator[](S32 i);
    int     batch_size();
    // main NN methods (model.h:85-91)
    Model &add(t4_layer fn, U32 n = 0, DU alpha = 0.0f, U16 *opt = nullptr);
    Model &forward(Tensor &input);
    Model &backprop();
    Model &backprop(Tensor &tgt);
    // loss functions (model.h:95-100)
    Tensor &onehot();
    Tensor &onehot(Tensor &t);
    Tensor &onehot_labels(const int32_t *labels_dev);        ///< Model::onehot(Dataset&) with device labels
    Model  &forward(Dataset &ds);                            ///< forward.cu:29-78 with a Dataset input: + onehot(ds) + hit (forward.cu:72-75), all on device
    int     step_graph(Dataset &ds, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd);   ///< commit + onehot + train step
    // the same, plus the loss read-back pipelined by one step: the loss of THIS step is copied to pinned host memory
    // asynchronously, *prev_loss receives the loss of the PREVIOUS call (NaN on the first); train_flush waits for the last one
    int     train_step(Dataset &ds, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd, DU *prev_loss);
    int     train_flush(DU *last_loss);
    int     hit(bool recalc = true);
    DU      loss(t4_loss op);
    DU      loss(t4_loss op, Tensor &tgt);
    int     loss_async(t4_loss op, Tensor &tgt, DU *loss_dev);
    // gradient descent (model.h:103-111)
    Model &grad_zero() { _iter = _hit = 0; return *this; }
    Model &grad_alloc(t4_optimizer op);
    Model &sgd(DU lr, DU b = 0.9f);
    Model &adam(DU lr, DU b1 = 0.9f, DU b2 = 0.999f);
    Model &adamw(DU lr, DU wd = 0.001f, DU b1 = 0.9f, DU b2 = 0.999f);
    // persistence (src/io/aio_model.cpp:16-235): `\\ tensorForth v4.0 model` header, one text line per layer (parameters glued to the
    // 7-character layer name exactly as AIO::_nsave_model writes them), a blank line, then per parametrised layer
    // `\n--- w.<name>\n` + raw FP32 of the weight (and `b.` bias; batchnorm: w only), closed by `\n---\n`.
    // load() reads the parameter sections into an already built model of the same architecture (AIO::nload, numel > 2 path).
    int    save(const char *fname, bool opt_state = false);   ///< opt_state: + the optimizer's moment arenas and step count behind the reference's sections (resume)
    int    load(const char *fname);
    static const char *nname(int fn);                        ///< model.cpp:17-21 (LAYER_OP, ntypes.h:47-51)
    int    arena(DU **G, DU **DG, int64_t *total);
    // data parallel: with a communicator attached, the optimizer calls (sgd/adam/adamw, also inside step_graph)
    // first SUM the gradient arena over the ranks — one fused exchange+optimizer kernel over NVLink peer memory (comm.cu);
    // `scal[0..nscal)` device floats (this rank's loss sum …) are summed over the ranks in the same exchange
    int    dp_attach(void *comm, DU *scal, int nscal);
    // this model holds shard `rank` of `world` equal shards of the global batch: dropout masks are drawn at the shard's global element offsets
    // (the masks of a single-device run of the whole batch), batch-norm statistics are SUM-all-reduced on `comm_stat` (a communicator of its
    // own, capacity >= 4 x bn_channels(); required before dp_attach when the model has batchnorm layers)
    int    dp_shard(int rank, int world, void *comm_stat);
    int    bn_channels();
    int    step_graph(Tensor &input, Tensor &tgt, t4_loss lop, DU *loss_dev, t4_optimizer op, DU lr, DU b1, DU b2, DU wd);
private:
    void _iconv(Tensor &in, U32 c, DU bias, U16 *opt, bool txn = false);
    void _ilinear(Tensor &in, U32 n, DU bias);
    void _iflatten(Tensor &in);
    void _isoftmax(Tensor &in);
    void _iactivate(Tensor &in, DU alpha);
    void _ipool(Tensor &in, U16 f);
    void _ibatchnorm(Tensor &in, DU m);
    void _iup(Tensor &in, U16 f, DU m);
    void _fstep(Tensor &in, Tensor &out);
    int  _ffused(size_t i, const DU *src = nullptr);
    int  _ffused_linear(size_t i);
    int  _bfused_head(Tensor &tgt, bool *skip_db);
    int  _bfused(int i);
    int  _fconv(Tensor &in, Tensor &out);
    int  _fdconv(Tensor &in, Tensor &out);
    int  _bdconv(Tensor &in, Tensor &out);
    int  _flinear(Tensor &in, Tensor &out);
    int  _factivate(Tensor &in, Tensor &out, t4_layer fn);
    int  _fpool(Tensor &in, Tensor &out, t4_layer fn);
    int  _fsoftmax(Tensor &in, Tensor &out);
    int  _flogsoftmax(Tensor &in, Tensor &out);
    int  _fbatchnorm(Tensor &in, Tensor &out);
    int  _fupsample(Tensor &in, Tensor &out);
    int  _bprep(Tensor &tgt);
    void _bstep(Tensor &in, Tensor &out, bool last_layer);
    int  _bconv(Tensor &in, Tensor &out);
    int  _blinear(Tensor &in, Tensor &out, bool skip_db = false, Tensor *xdup = nullptr, bool defer = false);
    int  _bactivate(Tensor &in, Tensor &out);
    int  _blinear_act(Tensor &in, Tensor &out, Tensor &act_in, bool skip_db);   ///< _blinear + the _bactivate of the activation in front (one GEMM epilogue)
    int  _bpool(Tensor &in, Tensor &out, t4_layer fn);
    int  _bupsample(Tensor &in, Tensor &out, t4_layer fn);
    int  _bbatchnorm(Tensor &in, Tensor &out);
    Model &_gradient(t4_optimizer op, DU lr, DU b1, DU b2, DU wd);
    void _RAND(Tensor &t, DU scale);
};

} // namespace t4
