#include "../glm.hpp"
