// Tidal potential at one cell: shared by the direct cell update (odis_kernels.cu) and the nonlinear cell update (odis_kernels_nl.cu).
#pragma once
#include "odis_kernels.cuh"

namespace odis {

// Per-cell trigonometric factors a potential needs (mesh.cpp:2132-2145), loaded up front.
struct TrigValues {
    double cosLat, sinLat, cosLon, sinLon, cos2Lat, sin2Lat, cos2Lon, sin2Lon, cosSq, sinSq;
};

// Tidal potential at one cell (tidalPotentials.cpp:80-172), same expression shapes.
__device__ __forceinline__ double tidal_potential(const Physics& p, const StepScalars& m, const TrigValues& v) {
    switch (p.potential) {
        case P_ECC:
            return p.factor * ((1. - 3. * v.sinSq) * m.cosM + v.cosSq * (3. * m.cosM * v.cos2Lon + 4. * m.sinM * v.sin2Lon));
        case P_OBLIQ:
            return p.factor * m.cosM * v.sin2Lat * v.cosLon;
        case P_OBLIQ_WEST:
            return 3 * p.factor * v.sinLat * v.cosLat * (v.cosLon * m.cosM - v.sinLon * m.sinM);
        case P_FULL:
            return p.factor * ((1 - 3 * v.sinSq) * m.cosM + v.cosSq * (3 * m.cosM * v.cos2Lon + 4 * m.sinM * v.sin2Lon)) +
                   p.factor2 * m.cosM * v.sin2Lat * v.cosLon;
        case P_FULL2: {
            const double ecc = p.ecc, obl = p.obl;
            double T1, T2, T3;
            T1 = 3. * ecc * (4. - 7. * obl * obl) * m.cosM + 6 * (obl * obl + ecc * ecc * (3 - 7 * obl * obl)) * m.cos2M;
            T1 += 3 * ecc * obl * obl * (7 * m.cos3M + 17 * ecc * m.cos4M);
            T1 *= -(1 - 3 * v.cos2Lat);
            T2 = (4 + 15 * ecc * ecc + 20 * ecc * m.cosM + 43 * ecc * ecc * m.cos2M) * v.cosLon;
            T2 += 2 * ecc * (4 + 25 * ecc * m.cosM) * m.sinM * v.sinLon;
            T2 *= 24 * obl * v.cosLat * v.sinLat * m.sinM;
            T3 = obl * obl * (2 + 3 * ecc * ecc + 6 * ecc * m.cosM + 9 * ecc * ecc * m.cos2M) * (m.cosM * v.cosLon + m.sinM * v.sinLon);
            T3 += -(obl * obl - 2) * ((6 * ecc * m.cosM + 17 * ecc * ecc * m.cos2M) * v.cos2Lon + 2 * ecc * (4 + 17 * ecc * m.cosM) * m.sinM * v.sin2Lon);
            T3 *= 6 * v.cosSq;
            return p.factor * (T1 + T2 + T3);
        }
        default:
            return 0.0;   // NONE leaves the (zero-initialised) potential untouched, tidalPotentials.cpp:283
    }
}


}  // namespace odis
