// ORACLE shim header — forwards to the single-file cv stand-in (see ../cvshim.h / ../../cvshim.h).
#pragma once
#include "cvshim.h"
