// convolve_kernel.cu — K3 placeholder (filled in below in this round).
#include "convolve.h"
namespace eb {}
