#pragma once
#include <boost/shim_core.hpp>
