#include "convolve.h"
namespace eb {}
