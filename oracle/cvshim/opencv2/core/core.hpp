// ORACLE shim forwarding header (test infrastructure only) — see cvshim.hpp
#pragma once
#include "../../cvshim.hpp"
