// placeholder until the lookahead kernels land
#include "ctx.h"
extern "C" void x264cu_lookahead_close_internal( x264cu_ctx *ctx ) { (void)ctx; }
