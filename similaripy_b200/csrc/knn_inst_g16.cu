// knn_inst_g16.cu -- the hot kernel with 16 lanes per B-row segment (see knn_inst.inc)
#define SPY_G 16
#include "knn_inst.inc"
