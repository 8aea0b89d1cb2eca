// stout.cu -- backward pass of the 4D stout layer (placeholder until the kernels land).
#include "gfb_internal.h"
namespace gfb {}
