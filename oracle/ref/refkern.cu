/*
 * refkern.cu — test driver that runs the REFERENCE's own CUDA kernels on raw buffers.
 *
 * TEST INFRASTRUCTURE ONLY (part of oracle/).  This file is ours; it is compiled against
 * the reference headers where they lie (/root/reference/src, -I on the nvcc line of
 * build_ref.sh) and linked with the reference's own kernel translation units
 * (src/t4math.cu, src/nn/nmath.cu).  No reference source is copied.
 *
 * It launches each kernel with the reference's own launch macros (FORK/FORK1/FORK3/FORK3T/
 * FORK4/FORK4P from src/t4base.h:129-159, src/t4math.h:117-123, src/nn/nmath.tcu:110-120) and,
 * where the reference's geometry lives in a .cu wrapper, restates that wrapper's launch
 * lines (citations inline).
 *
 * Protocol (little endian), request records until EOF on argv[1], responses to argv[2]:
 *   request : char op[16]; i32 ni; i32 iv[ni]; i32 nf; f32 fv[nf]; i32 na; { i64 len; f32 d[len] } x na
 *   response: i32 na; { i64 len; f32 d[len] } x na
 */
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <float.h>
#include "ten4_types.h"
#include "t4math.h"
#include "nn/nmath.h"

using namespace t4;
using namespace t4::nn;

struct Arr { long n; float *d; };           // device buffer
static std::vector<int>   iv;
static std::vector<float> fv;
static std::vector<Arr>   av;

static bool rd(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }

static bool read_req(FILE *f, char *op) {
    if (!rd(f, op, 16)) return false;
    int ni, nf, na;
    rd(f, &ni, 4); iv.resize(ni); if (ni) rd(f, iv.data(), 4 * ni);
    rd(f, &nf, 4); fv.resize(nf); if (nf) rd(f, fv.data(), 4 * nf);
    rd(f, &na, 4);
    for (auto &a : av) cudaFree(a.d);
    av.clear();
    for (int k = 0; k < na; k++) {
        long len; rd(f, &len, 8);
        std::vector<float> h(len);
        if (len) rd(f, h.data(), 4 * len);
        Arr a; a.n = len;
        cudaMalloc(&a.d, 4 * (len + 4));               // +scratch like MMU::talloc (numel+1)
        cudaMemset(a.d, 0, 4 * (len + 4));
        if (len) cudaMemcpy(a.d, h.data(), 4 * len, cudaMemcpyHostToDevice);
        av.push_back(a);
    }
    return true;
}
static void write_resp(FILE *f, std::vector<int> which) {
    int na = (int)which.size();
    fwrite(&na, 4, 1, f);
    for (int k : which) {
        std::vector<float> h(av[k].n);
        cudaMemcpy(h.data(), av[k].d, 4 * av[k].n, cudaMemcpyDeviceToHost);
        fwrite(&av[k].n, 8, 1, f);
        fwrite(h.data(), 4, av[k].n, f);
    }
}
#define TILE(v,t) (((v) + (t) - 1)/(t))
/* src/nn/forward.cu:116-124 (CONV macro) */
#define CONV(ks,s,p) do {                                       \
    constexpr int TS = (T4_DIM_SZ - (ks) + (s)) / (s);          \
    dim3 blk(T4_DIM_SZ, T4_DIM_SZ, 1);                          \
    dim3 grd(TILE(W0, TS), TILE(H0, TS), C0 * C1 * N);          \
    k_conv2d<TS, (ks), (s), (p)><<<grd,blk>>>(                  \
        I, O, F, B, H1, W1, H0, W0, C1, C0);                    \
    } while(0)
/* src/nn/backprop.cu:143-151 (DCONV macro) */
#define DCONV(ks,s,p) do {                                      \
    constexpr int TS = (T4_DIM_SZ - (ks) + (s)) / (s);          \
    dim3 blk(T4_DIM_SZ, T4_DIM_SZ, 1);                          \
    dim3 grd(TILE(W0, TS), TILE(H0, TS), C0 * C1 * N);          \
    k_dconv2d<TS, (ks), (s), (p)><<<grd,blk>>>(                 \
        I, dO, dX, F, dF, dB, H1, W1, H0, W0, C0, C1, train);   \
    } while(0)

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: refkern req.bin resp.bin\n"); return 2; }
    FILE *fi = fopen(argv[1], "rb"), *fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 2;
    char opn[17] = {0};
    while (read_req(fi, opn)) {
        std::string op(opn);
        if (op == "gemm") {                 /* Tensor::gemm1..4 per sample, src/mu/tensor.cu:125-201 */
            int variant = iv[0], tA = iv[1], tB = iv[2], M = iv[3], N = iv[4], K = iv[5], C = iv[6];
            float alpha = fv[0], beta = fv[1];
            float *A = av[0].d, *B = av[1].d, *O = av[2].d;
            switch (variant) {
            case 1: FORK3(k_gemm, M, N, C, A, B, O, alpha, beta, (bool)tA, (bool)tB, K); break;
            case 2: FORK3(k_gemm_claude, M, N, C, A, B, O, alpha, beta, (bool)tA, (bool)tB, K); break;
            case 4: FORK3T(k_gemm_tile_claude_x2, M, N, C, A, B, O, alpha, beta, (bool)tA, (bool)tB, K); break;
            default: FORK3T(k_gemm_tile_claude, M, N, C, A, B, O, alpha, beta, (bool)tA, (bool)tB, K); break;
            }
            write_resp(fo, {2});
        }
        else if (op == "map") {             /* Tensor::map, src/mu/tensor.cu:566-571 */
            long n = av[0].n;
            FORK(k_math, n, (math_op)iv[0], av[0].d, fv[0]);
            write_resp(fo, {0});
        }
        else if (op == "ts_op") {           /* src/mu/tensor.cu:17-23 */
            long n = av[0].n;
            FORK(k_ts_op, n, (math_op)iv[0], av[0].d, fv[0], av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "tt_op") {           /* src/mu/tensor.cu:50 */
            long n = av[0].n;
            FORK(k_tt_op, n, (math_op)iv[0], av[0].d, av[1].d, av[2].d);
            write_resp(fo, {2});
        }
        else if (op == "copy") {
            long n = av[0].n;
            FORK(k_copy, n, av[0].d, av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "transpose") {       /* src/mu/tensor.cu:210-219, one sample */
            int H = iv[0], W = iv[1], C = iv[2];
            FORK3(k_transpose, H, W, C, av[0].d, av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "identity") {
            int H = iv[0], W = iv[1], C = iv[2];
            FORK3(k_identity, H, W, C, av[0].d);
            write_resp(fo, {0});
        }
        else if (op == "sum") {             /* Tensor::sum, src/mu/tensor.cu:225-236 (GPU branch) */
            long n = av[0].n;
            FORK(k_sum, n, av[0].d, av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "nvar") {
            long n = av[0].n;
            FORK(k_nvar, n, av[0].d, fv[0], av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "max") {             /* Tensor::max/min, src/mu/tensor.cu:261-277 */
            long n = av[0].n;
            FORK(k_max, n, av[0].d, av[1].d, (bool)iv[0]);
            write_resp(fo, {1});
        }
        else if (op == "dot") {             /* src/mu/tensor.cu:61-72 */
            int K = iv[0], C = iv[1];
            FORK1(k_dot, C, 1, av[0].d, av[1].d, av[2].d, fv[0], fv[1], K, C);
            write_resp(fo, {2});
        }
        else if (op == "bce") {             /* src/mu/tensor.cu:307-312 */
            long n = av[0].n;
            FORK(k_bce, n, av[0].d, av[1].d, av[2].d);
            write_resp(fo, {2});
        }
        else if (op == "bias") {            /* src/nn/forward.cu:195 */
            int N = iv[0], E0 = iv[1];
            FORK3(k_bias, N, E0, 1, av[0].d, av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "dlinear_db") {      /* src/nn/backprop.cu:239 */
            int N = iv[0], E0 = iv[1];
            FORK3(k_dlinear_db, N, E0, 1, av[0].d, av[1].d);
            write_resp(fo, {1});
        }
        else if (op == "activate") {        /* src/nn/forward.cu:201-209 */
            long n = av[0].n;
            FORK(k_activate, n, (t4_layer)iv[0], av[0].d, av[1].d, av[2].d, fv[0]);
            write_resp(fo, {1, 2});
        }
        else if (op == "softmax") {         /* src/nn/forward.cu:231-243 */
            int N = iv[0], C = iv[1];
            if (C <= T4_DIM_SQ) { FORK2(k_softmax_small, N, C, av[0].d, av[1].d); }
            else                { FORK2(k_softmax, N, C, av[0].d, av[1].d); }
            write_resp(fo, {1});
        }
        else if (op == "conv2d") {          /* Model::_fconv, src/nn/forward.cu:126-155 */
            int N = iv[0], H1 = iv[1], W1 = iv[2], C1 = iv[3], H0 = iv[4], W0 = iv[5], C0 = iv[6];
            int KS = iv[7], S = iv[8], P = iv[9];
            float *I = av[0].d, *F = av[1].d, *B = av[2].d, *O = av[3].d;
            cudaMemset(O, 0, 4 * av[3].n);
            switch ((KS << 8) | (S << 4) | P) {
            case 0x110: CONV(1, 1, 0); break;
            case 0x311: CONV(3, 1, 1); break;
            case 0x421: CONV(4, 2, 1); break;
            case 0x512: CONV(5, 1, 2); break;
            default: fprintf(stderr, "conv cfg?\n"); return 3;
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) fprintf(stderr, "refkern conv2d launch: %s\n", cudaGetErrorString(e));
            GPU_CHK();
            write_resp(fo, {3});
        }
        else if (op == "dconv2d") {         /* Model::_bconv, src/nn/backprop.cu:153-191 */
            int N = iv[0], H1 = iv[1], W1 = iv[2], C1 = iv[3], H0 = iv[4], W0 = iv[5], C0 = iv[6];
            int KS = iv[7], S = iv[8], P = iv[9]; bool train = iv[10];
            float *I = av[0].d, *dO = av[1].d, *F = av[2].d, *dX = av[3].d, *dF = av[4].d, *dB = av[5].d;
            cudaMemset(dX, 0, 4 * av[3].n);
            switch ((KS << 8) | (S << 4) | P) {
            case 0x110: DCONV(1, 1, 0); break;
            case 0x311: DCONV(3, 1, 1); break;
            case 0x421: DCONV(4, 2, 1); break;
            case 0x512: DCONV(5, 1, 2); break;
            default: fprintf(stderr, "dconv cfg?\n"); return 3;
            }
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) fprintf(stderr, "refkern dconv2d launch: %s\n", cudaGetErrorString(e));
            GPU_CHK();
            write_resp(fo, {3, 4, 5});
        }
        else if (op == "pool" || op == "dpool") {   /* src/nn/forward.cu:212-228, backprop.cu:266-280 */
            t4_layer fn = (t4_layer)iv[0];
            int N = iv[1], H1 = iv[2], W1 = iv[3], H0 = iv[4], W0 = iv[5], C = iv[6], K = iv[7];
            if (op == "pool") {
                if (K == 2) { FORK4P(k_pool<2>, fn, av[0].d, av[1].d, H1, W1, H0, W0, C); }
                else        { FORK4P(k_pool<3>, fn, av[0].d, av[1].d, H1, W1, H0, W0, C); }
                write_resp(fo, {1});
            } else {
                if (K == 2) { FORK4P(k_dpool<2>, fn, av[0].d, av[1].d, H1, W1, H0, W0, C); }
                else        { FORK4P(k_dpool<3>, fn, av[0].d, av[1].d, H1, W1, H0, W0, C); }
                write_resp(fo, {0});
            }
        }
        else if (op == "batchnorm") {       /* Model::_fbatchnorm, src/nn/forward.cu:264-309 */
            int N = iv[0], H = iv[1], W = iv[2], C = iv[3];
            const int HW = H * W; const long NHW = (long)HW * N;
            float *I = av[0].d, *g = av[1].d, *b = av[2].d, *O = av[3].d, *XH = av[4].d;
            float *var = av[5].d, *avg = av[5].d + C;          // mtum[4]: var[C] | avg[C] | (s2)
            cudaMemset(var, 0, 2 * C * sizeof(float));
            { const int _b = std::max(32, std::min((int)HW, 1024)); const dim3 _g(C, N, 1);
              k_batchnorm_1<<<_g, _b>>>(I, avg, var, HW); GPU_CHK(); }
            { const int _b = std::max(32, std::min((int)C, 1024)); const int _g = (C + _b - 1) / _b;
              k_batchnorm_2<<<_g, _b>>>(avg, var, NHW, C); GPU_CHK(); }
            FORK4(k_batchnorm_3, 0, I, O, XH, g, b, avg, var, HW);
            write_resp(fo, {3, 4, 5});
        }
        else if (op == "dbatchnorm") {      /* Model::_bbatchnorm, src/nn/backprop.cu:312-370 */
            int N = iv[0], H = iv[1], W = iv[2], C = iv[3]; bool train = iv[4];
            const int HW = H * W; const long NHW = (long)HW * N;
            float *dO = av[0].d, *XH = av[1].d, *g = av[2].d, *dW = av[3].d, *dB = av[4].d;
            float *scr = av[5].d, *dX = av[6].d;
            float *var = scr, *s1 = scr + C, *s2 = scr + 2 * C;
            cudaMemset(s1, 0, C * 2 * sizeof(float));
            { const int nwarp = (T4_DIM_SQ + 31) >> 5; const int smem_sz = 2 * nwarp * sizeof(float);
              FORK4(k_dbatchnorm_1, smem_sz, dO, XH, s1, s2, HW); }
            { const int _b = std::max(32, std::min((int)C, 1024)); const int _g = ((int)C + _b - 1) / _b;
              k_dbatchnorm_2<<<_g, _b>>>(dW, dB, s1, s2, NHW, C, train); GPU_CHK(); }
            FORK4(k_dbatchnorm_3, 0, g, dO, XH, dX, s1, s2, var, HW);
            write_resp(fo, {6, 3, 4});
        }
        else if (op == "sgd") {             /* src/nn/gradient.cu:133-143 */
            long n = av[0].n;
            FORK(k_sgd, n, av[0].d, av[1].d, av[2].d, iv[0], fv[0], fv[1]);
            write_resp(fo, {0, 1, 2});
        }
        else if (op == "adam") {            /* src/nn/gradient.cu:145-157 */
            long n = av[0].n;
            FORK(k_adam, n, av[0].d, av[1].d, av[2].d, av[3].d, iv[0], fv[0], fv[1], fv[2]);
            write_resp(fo, {0, 1, 2, 3});
        }
        else if (op == "adamw") {           /* src/nn/gradient.cu:159-169 */
            long n = av[0].n;
            FORK(k_adamw, n, av[0].d, av[1].d, av[2].d, av[3].d, iv[0], fv[0], fv[1], fv[2], fv[3]);
            write_resp(fo, {0, 1, 2, 3});
        }
        else { fprintf(stderr, "refkern: unknown op '%s'\n", opn); return 4; }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { fprintf(stderr, "refkern %s: %s\n", opn, cudaGetErrorString(e)); return 5; }
    }
    fclose(fi); fclose(fo);
    return 0;
}
