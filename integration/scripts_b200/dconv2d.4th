\ dconv2d.4th — the conv-transpose layer (word `dconv2d`: 4x4, stride 2, padding 1) through the reference's VM on the new kernels (ten4_b200 only: the
\ reference wires the layer to its kernels with un-swapped tensors, src/nn/forward.cu:110, and produces no output).  Constant input 0.5, constant filter
\ 0.01, zero bias, 8 input channels: an interior output pixel collects 2 x 2 taps x 8 channels x 0.005 = 0.16, a corner pixel 1 tap = 0.04; the
\ backward pass with dY = 1 gives every input pixel its 4 x 4 taps x 6 output channels x 0.01 = 0.96 (interior) as dX.
0 trace
2 4 4 8 nn.model 0.0 6 dconv2d constant md0
md0
6 4 4 8 tensor ones 0.01 *=  0 nn.w=
drop
2 4 4 8 tensor ones 0.5 *= constant X
md0 X forward
." out max=" -1 n@ max . drop cr
." out min=" -1 n@ min . drop cr
." out sum=" -1 n@ sum . drop cr
2 8 8 6 tensor ones constant DY
DY backprop
." dx max=" 0 n@ max . drop cr
." db sum=" 0 nn.db sum . drop cr
drop
bye
