// Stand-in for <opencv2/features2d/features2d.hpp>: everything lives in cvshim.hpp (oracle/ref_shim, test infrastructure).
#include "../../cvshim.hpp"
