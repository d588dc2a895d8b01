// Tables that only the nonlinear branch (`advection; true`) reads: vertex geometry and the three sparse operators
// Mesh::operatorCurl (V x F), ::operatorRBFinterp (3N x F), ::operatorDirectionalSecondDeriv (2F x N), assembled as the
// reference does (/root/reference/src/mesh.cpp: CalcControlVolumeInterpMatrix :199-434, the vertex part of AssignFaces :910-1147,
// CalcVelocityTransformCoords :1335-1382, CalcMappingCoords :1384-1425, CalcRBFInterpMatrix :2263-2361, CalcRBFInterpMatrix2
// :2364-2719, CalcAdjacencyMatrix :2722-2806, CalcCurlOperatorCoeffs :3122-3175). The reference forms the operators as chains of
// Eigen sparse products; here each final coefficient is accumulated directly, in the order those products add their terms
// (row-wise, ascending inner index), and the small dense inverses use the same partial-pivot LU — so the CSR arrays can be compared
// entry for entry with the reference's.
#pragma once
#include <string>
#include <vector>

#include "odis_mesh.h"

namespace odis {

struct Csr {
    int n_rows = 0, n_cols = 0;
    std::vector<int> indptr, indices;
    std::vector<double> data;
};

struct NonlinearTables {
    std::vector<double> vertex_sinlat;     // [V]
    std::vector<double> vertex_area;       // [V]
    std::vector<int> vertex_faces;         // [V][3]
    std::vector<int> vertex_face_dir;      // [V][3]
    Csr curl, rbf_interp, directional_second_deriv;
};

int build_nonlinear_tables(const MeshTables& m, double radius, double rbf_eps, NonlinearTables& out, std::string& err);

}  // namespace odis
