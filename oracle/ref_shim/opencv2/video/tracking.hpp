#include "../../mini_cv.hpp"
