0 trace
512 constant N
N 28 28 1 nn.model 0.5 10 conv2d 2 maxpool relu flatten 100 linear relu 10 linear softmax constant md0
N 28 28 1 tensor rand 2 *= 1 -= constant X
N 1 10 1 tensor rand constant Y
: step ( M -- M ) X forward Y loss.ce drop Y backprop 0.001 nn.adam ;
: bench ( M n -- M ) clock >r for step next clock r> - . ;
md0 2 bench cr
19 bench cr
X forward Y loss.ce . cr
bye
