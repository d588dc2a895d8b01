// TEST INFRASTRUCTURE — not product code.
//
// The reference fixes the grid size at compile time (NODE_NUM / FACE_NUM / VERTEX_NUM from
// `constexpr int GRID_LVL = 6`, /root/reference/constants/gridConstants.h:17-46). This
// stand-in, found first on the include path, supplies the same three names from a
// -DODIS_REF_GRID_LVL=<L> build flag so one reference binary per grid level can be built
// without touching the reference tree. Level L has 10*4^(L-1)+2 cells.
#ifndef GRIDCONSTANTS_H
#define GRIDCONSTANTS_H
#include <math.h>
#ifndef ODIS_REF_GRID_LVL
#error "build with -DODIS_REF_GRID_LVL=<level>"
#endif
constexpr int GRID_LVL = ODIS_REF_GRID_LVL;
constexpr int odis_ref_cells(int lvl) { int s = 1; for (int i = 1; i < lvl; i++) s *= 2; return 10 * s * s + 2; }
const int NODE_NUM = odis_ref_cells(GRID_LVL);
const int FACE_NUM = 3 * NODE_NUM - 6;
const int VERTEX_NUM = 2 * NODE_NUM - 4;
#endif
