// oracle build shim: empty forward (reference includes it, uses nothing from it on the CPU engine path)
#include "opencv2/core/core.hpp"
