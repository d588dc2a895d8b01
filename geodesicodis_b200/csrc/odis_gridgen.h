// Synthetic icosahedral-bisection grids in the reference's grid_l<L>.txt conventions.
//
// The reference ships grid_l3..l6.txt only (input_files/); the larger files it names are
// missing blobs, so every configuration above 10,242 cells needs a generated stand-in. The
// conventions reproduced here were read off the shipped files and ReadMeshFile
// (/root/reference/src/mesh.cpp:4016-4101): cells 0-11 are the pentagons and carry a -1 sixth
// neighbour and a (-1,-1) sixth corner; neighbours run clockwise seen from outside; corner j is
// the circumcentre of (cell, neighbour j, neighbour j+1); longitudes lie in [0,360); new cells
// are appended level by level (hierarchical numbering, poor locality by construction).
#pragma once
#include <string>
#include "odis_mesh.h"
namespace odis {
// level L >= 2 gives 10*4^(L-1)+2 cells (constants/gridConstants.h:19-32). Coordinates are
// canonicalised through the "%.16f"-degree text form so that a grid written to a file and read
// back is bit-identical to the in-memory arrays.
int generate_icosahedral_grid(int level, GridFile& out, std::string& err);
}  // namespace odis
