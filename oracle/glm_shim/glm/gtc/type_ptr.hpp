// stand-in header, see glm.hpp
#pragma once
#include <glm/glm.hpp>
