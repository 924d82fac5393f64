// Stand-in for <opencv2/highgui/highgui.hpp>: everything lives in cvshim.hpp (oracle/ref_shim, test infrastructure).
#include "../../cvshim.hpp"
