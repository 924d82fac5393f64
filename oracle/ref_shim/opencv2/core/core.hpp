// Stand-in for <opencv2/core/core.hpp>: everything lives in cvshim.hpp (oracle/ref_shim, test infrastructure).
#include "../../cvshim.hpp"
