// knn_inst_g4.cu -- the hot kernel with 4 lanes per B-row segment (see knn_inst.inc)
#define SPY_G 4
#include "knn_inst.inc"
