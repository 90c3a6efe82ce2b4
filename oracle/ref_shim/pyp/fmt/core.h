#include "fmt.hpp"
