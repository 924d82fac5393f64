// Stand-in for <opencv/cv.h>: everything lives in cvshim.hpp (oracle/ref_shim, test infrastructure).
#include "../cvshim.hpp"
