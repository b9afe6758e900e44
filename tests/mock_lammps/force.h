#include "pair.h"
