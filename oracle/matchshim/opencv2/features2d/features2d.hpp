#include "../../mcv.h"
