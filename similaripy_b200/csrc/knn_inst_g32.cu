// knn_inst_g32.cu -- the hot kernel with 32 lanes per B-row segment (see knn_inst.inc)
#define SPY_G 32
#include "knn_inst.inc"
