/*
 * model_shim.cu — the launch layer of the reference's Model (src/nn/forward.cu, src/nn/backprop.cu) re-expressed on
 * libt4k.so, compiled AGAINST THE REFERENCE'S OWN HEADERS (src/nn/model.h, src/mu/tensor.h) into `ten4_b200`.
 * It defines exactly the Model methods those two reference files define (forward, _fstep, _f*, backprop, broadcast,
 * _bprep, _bstep, _b*); everything else of the reference (VM words, MMU, model.cpp, loss.cpp, gradient.cu, printing,
 * save/load) is compiled unmodified.  This is INTEGRATION.md §3 made real; t4host.cpp is the same code against this
 * repository's own class mirror.
 *
 * Semantics kept: layer i's tensor holds that layer's input; backprop overwrites activations with gradients in place;
 * error messages and the trace lines (`1 trace`) use the reference's format strings; host-synchronous at the end of
 * forward()/backprop() (the VM reads tensors on the host right after).  With tracing on, the per-layer path runs
 * (one call per layer, synchronised, so the per-layer Σ lines mean what they mean in the reference); with `0 trace`
 * the fused layer groups of include/t4k.h are used (same tensors written).
 */
#include <algorithm>
#include "nn/model.h"
#include "nn/nmath.h"
#include "mu/dataset.h"
#include "../include/t4k.h"

#if (T4_DO_OBJ && T4_DO_NN)
namespace t4::nn {

#define ST0 ((t4k_stream_t)0)                      /* legacy default stream, as the reference */
#define KCALL(call) do { int _rc = (call);                                                     \
    if (_rc > 0) { GPU_ERR((cudaError_t)_rc); }                                                 \
    else if (_rc < 0) { ERROR("%s -> %s\n", #call, t4k_strerror(_rc)); } } while (0)

static bool mask_act(t4_layer fn) { return fn == L_RELU || fn == L_TANH || fn == L_SELU || fn == L_LEAKYRL || fn == L_ELU; }

// ------------------------------------------------------------------------------------------ forward (forward.cu:29-78)
__HOST__ Model&
Model::forward(Tensor &input) {
    Tensor &n0 = (*this)[0];
    if (*_trace) input.show(true);
    if (input.numel != n0.numel) {
        ERROR("nn#forward dataset wrong shape[%d,%d,%d,%d] != model input[%d,%d,%d,%d]\n",
              input.N(), input.H(), input.W(), input.C(), n0.N(), n0.H(), n0.W(), n0.C());
        return *this;
    }
    const bool tr = *_trace != 0;
    NLOG("\nModel::forward starts trace=%d {", *_trace);
    DU t0 = System::clock(), t1 = t0, tt;
    int i = 0;
    const int n = (int)numel;
    // first layer group fused: the `n0 = input` copy rides in the same launch
    bool copied = false;
    if (!tr && input.data != n0.data && n >= 4) {
        Tensor &co = (*this)[1], &po = (*this)[2], &ao = (*this)[3];
        if (n0.grad_fn == L_CONV && co.grad_fn == L_MAXPOOL && co.stride[0] == 2 && po.grad_fn == L_RELU) {
            Tensor *fl = (ao.grad_fn == L_FLATTEN && n >= 5) ? &(*this)[4] : NULL;
            Tensor &f = *n0.grad[0], &b = *n0.grad[1];
            int rc = t4k_conv_pool_relu_fwd(input.data, f.data, b.data, n0.data, co.data, po.data, ao.data, po.grad[4]->data,
                                            fl ? fl->data : NULL, co.N(), n0.H(), n0.W(), n0.C(), co.H(), co.W(), co.C(),
                                            f.H(), n0.stride[0], n0.stride[2], ST0);
            if (rc != T4K_ENOSUP) { KCALL(rc); i = fl ? 4 : 3; copied = true; }
        }
    }
    if (!copied && input.data != n0.data) n0 = input;
    while (i < n - 1) {
        Tensor &in = (*this)[i], &out = (*this)[i + 1];
        if (tr) {
            GPU_CHK();
            INFO("\n%6.2f:%3d> %s [%2d,%2d,%2d,%2d] Σ/n=%6.2f p=%6.3f => out[%2d,%2d,%2d,%2d]",
                 (tt = System::clock()) - t1, i, nname(in.grad_fn), in.N(), in.H(), in.W(), in.C(),
                 in.sum() / in.N() / in.C(), in.xparm, out.N(), out.H(), out.W(), out.C());
            t1 = tt;
        }
        int adv = 0;
        if (!tr && i + 2 < n) {                                  // fused layer groups (same tensors written)
            Tensor &o2 = (*this)[i + 2];
            const t4_layer f1 = in.grad_fn, f2 = out.grad_fn;
            if (f1 == L_CONV && f2 == L_MAXPOOL && out.stride[0] == 2 && o2.grad_fn == L_RELU && i + 3 < n) {
                Tensor &ao = (*this)[i + 3];
                Tensor *fl = (ao.grad_fn == L_FLATTEN && i + 4 < n) ? &(*this)[i + 4] : NULL;
                Tensor &f = *in.grad[0], &b = *in.grad[1];
                int rc = t4k_conv_pool_relu_fwd(in.data, f.data, b.data, NULL, out.data, o2.data, ao.data, o2.grad[4]->data,
                                                fl ? fl->data : NULL, out.N(), in.H(), in.W(), in.C(), out.H(), out.W(), out.C(),
                                                f.H(), in.stride[0], in.stride[2], ST0);
                if (rc != T4K_ENOSUP) { KCALL(rc); adv = fl ? 4 : 3; }
            }
            else if (f1 == L_LINEAR && (f2 == L_SOFTMAX || mask_act(f2) || f2 == L_SIGMOID)) {
                const int N = (int)out.N(), E0 = (int)out.HWC(), E1 = (int)in.HWC();
                int rc = (f2 == L_SOFTMAX)
                    ? t4k_mlp_head_fwd(in.data, in.grad[0]->data, in.grad[1]->data, out.data, o2.data, N, E0, E1, ST0)
                    : t4k_linear_act_fwd(f2, in.data, in.grad[0]->data, in.grad[1]->data, out.data, o2.data, out.grad[4]->data,
                                         out.xparm, N, E0, E1, ST0);
                if (rc != T4K_ENOSUP) { KCALL(rc); adv = 2; }
            }
        }
        if (adv) { i += adv; continue; }
        _fstep(in, out);
        if (tr) {
            GPU_CHK();
            if (_check_nan(out)) {
                ERROR("nn#forward Nan in %s\n", nname(in.grad_fn));
                INFO("in=");  in.show(true);
                INFO("out="); out.show(true);
                this->err = 1;
                break;
            }
            if (*_trace > 1) out.show(true);
        }
        i++;
    }
    GPU_CHK();                                                   // host-synchronous, once
    if (input.is_dataset()) {
        onehot((Dataset&)input);
        _hit = hit(true);
    }
    NLOG("\n} Model::forward %5.2f ms\n", System::clock() - t0);
    return *this;
}

__HOST__ void
Model::_fstep(Tensor &in, Tensor &out) {                          // forward.cu:83-113
    t4_layer fn = in.grad_fn;
    switch (fn) {
    case L_CONV:    _fconv(in, out);         break;
    case L_LINEAR:  _flinear(in, out);       break;
    case L_FLATTEN: KCALL(t4k_copy(in.data, out.data, in.numel, ST0)); break;
    case L_RELU: case L_TANH: case L_SIGMOID: case L_SELU: case L_LEAKYRL:
    case L_ELU:     _factivate(in, out, fn); break;
    case L_DROPOUT: {
        Tensor &t = *in.grad[4];
        System::rand(t.data, t.numel, UNIFORM);                   // fresh mask every forward (forward.cu:98-102)
        _factivate(in, out, fn);
    } break;
    case L_SOFTMAX: _fsoftmax(in, out);      break;
    case L_LOGSMAX: _flogsoftmax(in, out);   break;
    case L_AVGPOOL: case L_MAXPOOL:
    case L_MINPOOL: _fpool(in, out, fn);     break;
    case L_BATCHNM: _fbatchnorm(in, out);    break;
    case L_USAMPLE: _fupsample(in, out);     break;
    case L_DCONV: {                                               // forward.cu:110: the convolution's kernels with swapped roles (t4k_dconv2d_fwd);
        Tensor &f = *in.grad[0], &b = *in.grad[1];                // the reference's filter T4(C1,K,K,C0) is read as [C0][K][K][C1] (same numel)
        int rc = t4k_dconv2d_fwd(in.data, f.data, b.data, out.data, out.N(), in.H(), in.W(), in.C(), out.H(), out.W(), out.C(),
                                 f.H(), in.stride[0], in.stride[2], ST0);
        if (rc == T4K_ENOSUP) ERROR("nn#fconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]);
        else KCALL(rc);
    } break;
    default: ERROR("nn#fstep layer=%d not supported\n", fn);
    }
}
__HOST__ int
Model::_fconv(Tensor &in, Tensor &out) {                          // forward.cu:126-155
    Tensor &f = *in.grad[0], &b = *in.grad[1];
    int rc = t4k_conv2d_fwd(in.data, f.data, b.data, out.data, out.N(), in.H(), in.W(), in.C(), out.H(), out.W(), out.C(),
                            f.H(), in.stride[0], in.stride[2], ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#fconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCALL(rc);
    return 0;
}
__HOST__ int
Model::_flinear(Tensor &in, Tensor &out) {                        // forward.cu:158-198 (GEMM + bias)
    if (*_trace > 1) {                                             // level-2 dumps, same places as the reference
        GPU_CHK();
        _dump_w("w", *in.grad[0], in.grad[0]->numel < T4_DIM_SQ);
        _dump_b("b", *in.grad[1]); INFO("\n");
    }
    KCALL(t4k_linear_fwd(in.data, in.grad[0]->data, in.grad[1]->data, out.data, out.N(), (int)out.HWC(), (int)in.HWC(), ST0));
    return 0;
}
__HOST__ int
Model::_factivate(Tensor &in, Tensor &out, t4_layer fn) {         // forward.cu:201-209
    KCALL(t4k_activate_fwd(fn, in.data, out.data, in.grad[4]->data, in.xparm, in.numel, ST0));
    if (train && *_trace > 1) { GPU_CHK(); _dump_f("msk", *in.grad[4]); }
    return 0;
}
__HOST__ int
Model::_fpool(Tensor &in, Tensor &out, t4_layer fn) {             // forward.cu:212-228
    int rc = t4k_pool_fwd(fn, in.data, out.data, out.N(), in.H(), in.W(), out.H(), out.W(), out.C(), in.stride[0], ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#fpool kernel_size=%d not supported\n", in.stride[0]); return -1; }
    KCALL(rc);
    return 0;
}
__HOST__ int
Model::_fsoftmax(Tensor &in, Tensor &out) {                       // forward.cu:231-243
    KCALL(t4k_softmax_fwd(in.data, out.data, in.N(), (int)in.HWC(), ST0));
    return 0;
}
__HOST__ int
Model::_flogsoftmax(Tensor &in, Tensor &out) {                    // forward.cu:246-259 (as coded)
    KCALL(t4k_logsoftmax_fwd(in.data, out.data, in.N(), (int)in.HWC(), ST0));
    return 0;
}
__HOST__ int
Model::_fbatchnorm(Tensor &in, Tensor &out) {                     // forward.cu:264-309
    KCALL(t4k_batchnorm_fwd(in.data, out.data, in.grad[4]->data, in.grad[0]->data, in.grad[1]->data, in.mtum[4]->data,
                            out.N(), out.H() * out.W(), out.C(), ST0));
    if (*_trace > 1) {
        GPU_CHK();
        _dump_b("w", *in.grad[0]); _dump_b("b", *in.grad[1]);
        INFO("\n    xht="); in.grad[4]->show();
    }
    return 0;
}
__HOST__ int
Model::_fupsample(Tensor &in, Tensor &out) {                      // forward.cu:314-329
    int rc = t4k_pool_bwd(T4K_L_USAMPLE, out.data, in.data, in.N(), out.H(), out.W(), in.H(), in.W(), in.C(), in.stride[0], ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#fupsample size=%d not supported\n", in.stride[0]); return -1; }
    KCALL(rc);
    return 0;
}

// ------------------------------------------------------------------------------------------ backprop (backprop.cu:16-140)
__HOST__ Model&
Model::broadcast(Tensor &tgt) {                                   // [N,1] target → [N,HWC] cached one-hot
    Tensor &out = (*this)[-1];
    U64 HWC = out.HWC();
    U32 N   = out.N();
    if (!_hot) _hot = &T4(N, 1, HWC, 1);
    GPU_CHK();
    for (U32 n = 0; n < N; n++) {
        DU  v = tgt.data[n];
        DU *h = _hot->slice(n);
        for (U64 i = 0; i < HWC; i++) h[i] = v;
    }
    return *this;
}
__HOST__ Model&
Model::backprop() {
    if (_hot) return backprop(*_hot);
    ERROR("nn#backprop missing onehot vector?\n");
    return *this;
}
__HOST__ Model&
Model::backprop(Tensor &tgt) {
    const bool tr = *_trace != 0;
    const int n = (int)numel;
    int i = n - 2, j = 0;
    bool skip_db = false, head = false;
    Tensor &outl = (*this)[-1];
    // classifier head [linear →] activation → small linear → softmax: one launch (see t4k_mlp_head_bwd)
    if (!tr && n >= 4 && outl.numel == tgt.numel) {
        Tensor &yl = (*this)[n - 2], &x2 = (*this)[n - 3];
        const int N = (int)outl.N(), E0 = (int)outl.HWC(), E1 = (int)x2.HWC();
        if (yl.grad_fn == L_SOFTMAX && x2.grad_fn == L_LINEAR && E0 <= 32 && E1 <= 128) {
            Tensor *act = (n >= 5 && (mask_act((*this)[n - 4].grad_fn) || (*this)[n - 4].grad_fn == L_DROPOUT)) ? &(*this)[n - 4] : NULL;
            const int prev = act ? n - 5 : n - 4;
            Tensor *lin1 = (prev >= 0 && (*this)[prev].grad_fn == L_LINEAR && train) ? &(*this)[prev] : NULL;
            int rc = t4k_mlp_head_bwd(outl.data, tgt.data, yl.data, x2.data, act ? act->grad[4]->data : NULL, act ? act->data : NULL,
                                      x2.grad[0]->data, x2.grad[2]->data, x2.grad[3]->data, lin1 ? lin1->grad[3]->data : NULL,
                                      N, E0, E1, train, ST0);
            if (rc != T4K_ENOSUP) { KCALL(rc); head = true; skip_db = lin1 != NULL; i = prev; j = 1; }
        }
    }
    if (!head && _bprep(tgt)) return *this;

    NLOG("\nModel::backprop starts trace=%d train=%d {", *_trace, train);
    DU t0 = System::clock(), t1 = t0, tt;
    for (; i >= 0; j++) {
        Tensor &in = (*this)[i], &out = (*this)[i + 1];
        if (tr) {
            GPU_CHK();
            INFO("\n%6.2f:%3d> %s [%2d,%2d,%2d,%2d] p=%6.3f <= out'Σ/n=%6.2f [%2d,%2d,%2d,%2d]",
                 (tt = System::clock()) - t1, i, nname(in.grad_fn), in.N(), in.H(), in.W(), in.C(), in.xparm,
                 out.sum() / out.N() / out.C(), out.N(), out.H(), out.W(), out.C());
            t1 = tt;
        }
        const t4_layer fn = in.grad_fn;
        if (!tr && j > 0 && (fn == L_FLATTEN || fn == L_RELU)) {   // conv → maxpool(2) → relu (→ flatten), backward, one launch
            const bool flat = fn == L_FLATTEN;
            const int ir = flat ? i - 1 : i;
            if (ir >= 2) {
                Tensor &ci = (*this)[ir - 2], &co = (*this)[ir - 1], &po = (*this)[ir], &ao = (*this)[ir + 1];
                if (po.grad_fn == L_RELU && co.grad_fn == L_MAXPOOL && co.stride[0] == 2 && ci.grad_fn == L_CONV) {
                    Tensor &f = *ci.grad[0], &df = *ci.grad[2], &db = *ci.grad[3], &dx = *ci.grad[4];
                    int rc = t4k_conv_pool_relu_bwd(out.data, ao.data, po.grad[4]->data, po.data, co.data, ci.data, dx.data, f.data,
                                                    df.data, db.data, ci.N(), ci.H(), ci.W(), ci.C(), co.H(), co.W(), co.C(),
                                                    f.H(), ci.stride[0], ci.stride[2], train, ST0);
                    if (rc != T4K_ENOSUP) { KCALL(rc); i -= flat ? 4 : 3; continue; }
                }
            }
        }
        if (skip_db && fn == L_LINEAR) {                           // dB already accumulated by the head kernel
            KCALL(t4k_linear_bwd_ex(in.data, in.grad[0]->data, out.data, in.data, in.grad[2]->data, in.grad[3]->data,
                                    in.N(), (int)out.HWC(), (int)in.HWC(), train, 1, ST0));
            skip_db = false;
        }
        else _bstep(in, out, j == 0);
        if (tr) {
            GPU_CHK();
            if (_check_nan(in)) {
                ERROR("nn#backprop Nan %s\n", nname(in.grad_fn));
                in.show(); out.show();
                this->err = 1;
                break;
            }
            if (*_trace > 1) in.show(true);
        }
        i--;
    }
    GPU_CHK();                                                    // host-synchronous, once
    NLOG("\n} Model::backprop %5.2f ms\n", System::clock() - t0);
    return *this;
}
__HOST__ int
Model::_bprep(Tensor &tgt) {                                      // backprop.cu:76-109
    Tensor &out = (*this)[-1];
    if (out.numel != tgt.numel) {
        ERROR("Model#bprep: Onehot wrong shape[%d,%d,%d,%d] != [%d,%d,%d,%d], numel=%ld,%ld ",
              tgt.N(), tgt.H(), tgt.W(), tgt.C(), out.N(), out.H(), out.W(), out.C(), tgt.numel, out.numel);
        return 1;
    }
    NLOG("Model::bprep input(onehot) numel=%ld OK {\n", tgt.numel);
    switch ((*this)[-2].grad_fn) {
    case L_LINEAR: case L_SIGMOID: case L_SOFTMAX:
    case L_LOGSMAX: KCALL(t4k_tt_op(T4K_SUB, out.data, tgt.data, out.data, out.numel, 1, 1, ST0)); break;   // p - y, not divided by N
    default:        KCALL(t4k_copy(tgt.data, out.data, tgt.numel, ST0)); break;
    }
    if (*_trace) { GPU_CHK(); out.show(true); }
    NLOG("}\n");
    return 0;
}
__HOST__ void
Model::_bstep(Tensor &in, Tensor &out, bool last_layer) {         // backprop.cu:112-140
    t4_layer fn = in.grad_fn;
    switch (fn) {
    case L_CONV:    _bconv(in, out);         break;
    case L_LINEAR:
        if (last_layer) KCALL(t4k_copy(out.data, in.data, out.numel, ST0));        // linear + MSE
        else            _blinear(in, out);
        break;
    case L_FLATTEN: KCALL(t4k_copy(out.data, in.data, out.numel, ST0)); break;
    case L_RELU: case L_TANH: case L_SELU: case L_LEAKYRL: case L_ELU:
    case L_DROPOUT: _bactivate(in, out);     break;
    case L_SIGMOID: case L_SOFTMAX:
    case L_LOGSMAX: KCALL(t4k_copy(out.data, in.data, out.numel, ST0)); break;     // pass-through, hidden sigmoids too (backprop.cu:129-131)
    case L_MAXPOOL: case L_AVGPOOL:
    case L_MINPOOL: _bpool(in, out, fn);     break;
    case L_BATCHNM: _bbatchnorm(in, out);    break;
    case L_USAMPLE: _bupsample(in, out, fn); break;
    case L_DCONV: {                                               // backprop.cu:137
        Tensor &f = *in.grad[0], &df = *in.grad[2], &db = *in.grad[3], &dx = *in.grad[4];
        int rc = t4k_dconv2d_bwd(in.data, out.data, f.data, dx.data, df.data, db.data, in.N(), in.H(), in.W(), in.C(),
                                 out.H(), out.W(), out.C(), f.H(), in.stride[0], in.stride[2], train, ST0);
        if (rc == T4K_ENOSUP) ERROR("nn#bconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]);
        else { KCALL(rc); KCALL(t4k_copy(dx.data, in.data, dx.numel, ST0)); }     // x = dX (overwrite), as Model::_bconv
    } break;
    default: ERROR("nn#bstep layer=%d not supported\n", fn);
    }
}
__HOST__ int
Model::_bconv(Tensor &in, Tensor &out) {                          // backprop.cu:153-191
    Tensor &f = *in.grad[0], &df = *in.grad[2], &db = *in.grad[3], &dx = *in.grad[4];
    if (*_trace > 1) {
        GPU_CHK();
        _dump_b("before b", *in.grad[1]); _dump_f("before f", f);
        _dump_b("before db", db); _dump_f("before df", df); INFO("\n");
    }
    int rc = t4k_conv2d_bwd(in.data, out.data, f.data, dx.data, df.data, db.data, in.N(), in.H(), in.W(), in.C(),
                            out.H(), out.W(), out.C(), f.H(), in.stride[0], in.stride[2], train, ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#bconv kernel_size=%d stride=%d padding=%d not supported\n", f.H(), in.stride[0], in.stride[2]); return -1; }
    KCALL(rc);
    KCALL(t4k_copy(dx.data, in.data, dx.numel, ST0));             // in = dx
    if (*_trace > 1) { GPU_CHK(); _dump_b("after db", db); _dump_f("after df", df); INFO("\n"); }
    return 0;
}
__HOST__ int
Model::_blinear(Tensor &in, Tensor &out) {                        // backprop.cu:194-254
    Tensor &dw = *in.grad[2], &db = *in.grad[3];
    if (train && *_trace > 1) {
        GPU_CHK();
        _dump_b("before db", db); _dump_w("before dw", dw, dw.numel < T4_DIM_SQ); INFO("\n");
    }
    KCALL(t4k_linear_bwd(in.data, in.grad[0]->data, out.data, in.data, in.grad[2]->data, in.grad[3]->data,
                         in.N(), (int)out.HWC(), (int)in.HWC(), train, ST0));
    if (train && *_trace > 1) {
        GPU_CHK();
        _dump_b("after db", db); _dump_w("after dw", dw, dw.numel < T4_DIM_SQ); INFO("\n");
    }
    return 0;
}
__HOST__ int
Model::_bactivate(Tensor &in, Tensor &out) {                      // backprop.cu:257-263
    if (train && *_trace > 1) { GPU_CHK(); _dump_f("msk", *in.grad[4]); }
    KCALL(t4k_activate_bwd(out.data, in.grad[4]->data, in.data, in.numel, ST0));
    return 0;
}
__HOST__ int
Model::_bpool(Tensor &in, Tensor &out, t4_layer fn) {             // backprop.cu:266-280
    int rc = t4k_pool_bwd(fn, in.data, out.data, out.N(), in.H(), in.W(), out.H(), out.W(), out.C(), in.stride[0], ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#bpool kernel_size=%d not supported\n", in.stride[0]); return -1; }
    KCALL(rc);
    return 0;
}
__HOST__ int
Model::_bupsample(Tensor &in, Tensor &out, t4_layer fn) {         // backprop.cu:285-300
    int rc = t4k_pool_fwd(T4K_L_USAMPLE, out.data, in.data, in.N(), out.H(), out.W(), in.H(), in.W(), in.C(), in.stride[0], ST0);
    if (rc == T4K_ENOSUP) { ERROR("nn#bupsample size=%d not supported\n", in.stride[0]); return -1; }
    KCALL(rc);
    return 0;
}
__HOST__ int
Model::_bbatchnorm(Tensor &in, Tensor &out) {                     // backprop.cu:312-370
    KCALL(t4k_batchnorm_bwd(out.data, in.grad[4]->data, in.data, in.grad[0]->data, in.grad[2]->data, in.grad[3]->data,
                            in.mtum[4]->data, in.N(), in.H() * in.W(), in.C(), train, ST0));
    if (train && *_trace > 1) {                                    // (the reference also prints its intermediate per-stage sums; one fused pass here)
        GPU_CHK();
        _dump_b("db=sum_dout     ", *in.grad[3]); _dump_b("dw-sum_dout_xhat", *in.grad[2]); INFO("\n");
    }
    return 0;
}

} // namespace t4::nn
#endif
