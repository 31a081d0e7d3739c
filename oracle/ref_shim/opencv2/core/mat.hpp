/* oracle shim: forwards to cvshim.hpp (test infrastructure, see that file). */
#pragma once
#include "../../cvshim.hpp"
