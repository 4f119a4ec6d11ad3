\ cnn_train_lr1e-3.4th — deterministic MNIST-shaped CNN training run (no rand, no dataset files) for the side-by-side
\ check of `ten4` (reference build, oracle/_ref) and `ten4_b200` (reference VM on libt4k.so, integration/_build).
\ Model of examples/t4_40a.4th:10-13 at N=64.  Weights, bias and input are logistic-map sequences built from
\ tensor words whose arithmetic is single IEEE ops in both builds (copy, ts/tt mul, add), so both builds start
\ from bit-identical tensors.
0 trace
64 constant N
: lg ( T -- T' ) copy -1 *= 1 += *= 4 *= ;                 \ x <- 4 x (1 - x)
: chaos ( T -- T' ) gradfill 0.8 *= 0.1 += 19 for lg next 0.5 -= ;   \ x0 = 0.1 + 0.8 j/n, 20 iterations, centred
N 28 28 1 nn.model 0.5 10 conv2d 2 maxpool relu flatten 0.0 100 linear relu 0.0 10 linear softmax constant md0
md0
1 3 3 10 tensor chaos 1.6 *=     0 nn.w=
10 vector chaos 0.2 *=           0 nn.b=
100 1960 matrix chaos 0.05 *=    4 nn.w=
100 vector chaos 0.1 *=          4 nn.b=
10 100 matrix chaos 0.2 *=       6 nn.w=
10 vector chaos 0.1 *=           6 nn.b=
drop
N 28 28 1 tensor chaos 2 *= constant X
N 1 10 1 tensor zeros constant Y
: hot ( -- ) N 1 - for Y 1 i 10 * i 7 * 3 + 10 mod + t! drop next ;   \ label(i) = (7 i + 3) mod 10
hot
." X sum=" X sum . drop cr
." labels sum=" Y sum . drop cr
." w0=" md0 0 nn.w . drop cr
md0 X forward
." out sum=" -1 n@ sum . drop cr
." out=" -1 n@ . cr
Y loss.ce ." loss0=" . cr
Y backprop
." dw0=" 0 nn.dw . cr
." db0=" 0 nn.db . cr
." dw4 sum=" 4 nn.dw sum . ."  norm=" norm . drop cr
." db4=" 4 nn.db . cr
." dw6 norm=" 6 nn.dw norm . drop cr
." db6=" 6 nn.db . cr
0.001 nn.adam
." w0'=" 0 nn.w . cr
\ nine more steps at the BASELINE learning rate (chaotic regime, see cnn_parity.4th)
: step ( M -- M ) X forward Y backprop 0.001 nn.adam ;
: run ( M n -- M ) for step X forward Y loss.ce ." loss=" . cr next ;
9 run
." w6 norm=" 6 nn.w norm . drop cr
." b4=" 4 nn.b . cr
drop
bye
