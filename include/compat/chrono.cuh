// forwarder: the reference declares this header separately; see simplex_compat.hpp
#pragma once
#include "simplex_compat.hpp"
