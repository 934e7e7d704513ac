#pragma once
#include <boost/geometry/geometry.hpp>
