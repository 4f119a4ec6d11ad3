\ big_arena.4th — SURVEY.md §8f row 3: tensors beyond the reference's 2 GiB object store, driven from Forth on the new kernels (ten4_b200 only:
\ the reference's `ten4` cannot allocate them).  (1) 4096x4096 GEMM through the `@` word (BASELINE config 2): ones @ ones = 4096 everywhere;
\ (2) a conv2d layer at N=1024, 56x56x64 -> 64 (BASELINE config 5's layer, 822 MB per tensor, 3.3 GB for the model) forward + backprop:
\ with a constant input, a constant filter and zero bias every interior output equals 9 * 64 * x * f.
0 trace
4096 4096 matrix ones
4096 4096 matrix ones
@ ." gemm sum/4096^3=" sum 4096 / 4096 / 4096 / . drop drop drop cr
1024 constant N
N 56 56 64 nn.model 0.0 64 conv2d constant md0
md0
64 3 3 64 tensor ones 0.001 *=  0 nn.w=
drop
N 56 56 64 tensor ones 0.5 *= constant X
md0 X forward
." conv out max=" -1 n@ max . drop cr
." conv out min=" -1 n@ min . drop cr
N 56 56 64 tensor ones constant DY
DY backprop
." dw sum/1e6=" 0 nn.dw sum 1000000 / . drop cr
drop
bye
