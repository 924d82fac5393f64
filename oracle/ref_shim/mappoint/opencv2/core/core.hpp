#include "../../../matcher/cv_float.hpp"
