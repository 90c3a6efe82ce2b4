#include "../../thread.hpp"
