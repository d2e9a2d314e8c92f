#include "dump_image.h"
#include "dump_sites.h"
#include "dump_text.h"
#include "dump_vtk.h"
