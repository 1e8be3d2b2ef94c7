// knn_inst_g8.cu -- the hot kernel with 8 lanes per B-row segment (see knn_inst.inc)
#define SPY_G 8
#include "knn_inst.inc"
