// sg2d_multi_hi.cu -- the half-window 5..8 instantiations of the multi-output 2D kernel (see sg2d_multi.cu).
#define SG2D_MULTI_HI 1
#include "sg2d_multi.cu"
