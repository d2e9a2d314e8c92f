#include "pair_lj_cut.h"
