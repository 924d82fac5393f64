// Stand-in for <opencv2/core/cuda.hpp>: everything lives in cvshim.hpp (oracle/ref_shim, test infrastructure).
#include "../../cvshim.hpp"
