// Material.h:3 includes "Maths.h"; the file on disk is maths.h (Windows tree).
#pragma once
#include "maths.h"
