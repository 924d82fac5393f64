#include "../../cv_float.hpp"
