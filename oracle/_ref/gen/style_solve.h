#include "solve_group.h"
#include "solve_linear.h"
#include "solve_tree.h"
