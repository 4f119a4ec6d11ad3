\ mlp_bn_parity.4th — deterministic MLP with batchnorm / tanh / leakyrelu / sigmoid + MSE + SGD (momentum), N=32;
\ the layer kinds the CNN script does not touch.  Same construction as cnn_parity.4th.
0 trace
32 constant N
: lg ( T -- T' ) copy -1 *= 1 += *= 4 *= ;
: chaos ( T -- T' ) gradfill 0.8 *= 0.1 += 19 for lg next 0.5 -= ;
N 8 8 2 nn.model flatten 0.0 48 linear batchnorm tanh 0.0 24 linear 0.1 leakyrelu 0.0 4 linear sigmoid constant md1
md1
48 128 matrix chaos 0.3 *=   1 nn.w=
48 vector chaos 0.2 *=       1 nn.b=
24 48 matrix chaos 0.4 *=    4 nn.w=
24 vector chaos 0.2 *=       4 nn.b=
4 24 matrix chaos 0.5 *=     6 nn.w=
4 vector chaos 0.2 *=        6 nn.b=
drop
N 8 8 2 tensor chaos 2 *= constant X
N 1 4 1 tensor chaos 0.5 += constant T
md1 network
X forward
." out=" -1 n@ . cr
T loss.mse ." loss0=" . cr
T backprop
." dw6=" 6 nn.dw . cr
." dw4 norm=" 4 nn.dw norm . drop cr
." dw1 norm=" 1 nn.dw norm . drop cr
." db1=" 1 nn.db . cr
0.01 0.9 nn.sgd
: step ( M -- M ) X forward T backprop 0.01 0.9 nn.sgd ;
: run ( M n -- M ) for step X forward T loss.mse ." loss=" . cr next ;
9 run
." w1 norm=" 1 nn.w norm . drop cr
drop
bye
