// Spherical / stereographic geometry primitives used while building the C-grid tables.
//
// Each routine evaluates the same closed-form expression, in the same floating-point
// operation order, as the reference helper it replaces (cited per function, all in
// /root/reference/include/mathRoutines.h) so that the FP64 tables derived from them are
// bit-identical to the reference's when both are compiled without FMA contraction.
// Points are (lat, lon) in radians.
#pragma once
#include <cmath>

namespace odis {

// mathRoutines.h:10-11
constexpr double kPi = 3.1415926535897932384626433832795028841971693993751058;
constexpr double kRadPerDeg = kPi / 180.0;

struct LatLon {
    double lat, lon;
};
struct Vec2 {
    double x, y;
};
struct Vec3 {
    double x, y, z;
};

// Stereographic map factor of point p seen from map centre c.  mathRoutines.h:30-33
inline double map_factor(LatLon c, LatLon p) {
    return 2.0 / (1.0 + std::sin(c.lat) * std::sin(p.lat) + std::cos(c.lat) * std::cos(p.lat) * std::cos(p.lon - c.lon));
}

// Stereographic projection of p about centre c on a sphere of radius r; returns the map
// factor through m.  mathRoutines.h:61-68
inline Vec2 map_project(LatLon c, LatLon p, double r, double* m_out = nullptr) {
    const double m = map_factor(c, p);
    Vec2 q;
    q.x = m * r * (std::cos(p.lat) * std::sin(p.lon - c.lon));
    q.y = m * r * (std::sin(p.lat) * std::cos(c.lat) - std::cos(p.lat) * std::sin(c.lat) * std::cos(p.lon - c.lon));
    if (m_out) *m_out = m;
    return q;
}

// r * acos(...) great-circle length.  mathRoutines.h:130-133
inline double arc_length_acos(LatLon a, LatLon b, double r) {
    return r * std::acos(std::sin(a.lat) * std::sin(b.lat) + std::cos(a.lat) * std::cos(b.lat) * std::cos(b.lon - a.lon));
}

// Unit vector of a (lat,lon) point expressed through colatitude, as sph2cart(1, pi/2-lat, lon).
// mathRoutines.h:19-24 as called from :144-145,254-255
inline Vec3 unit_from_colat(LatLon p) {
    const double colat = kPi * 0.5 - p.lat;
    Vec3 v;
    v.x = 1.0 * std::sin(colat) * std::cos(p.lon);
    v.y = 1.0 * std::sin(colat) * std::sin(p.lon);
    v.z = 1.0 * std::cos(colat);
    return v;
}

// atan2(|a x b|, a.b) angular distance (radians).  mathRoutines.h:139-157
inline double arc_angle_atan2(LatLon a, LatLon b) {
    const Vec3 v1 = unit_from_colat(a), v2 = unit_from_colat(b);
    const double cx = v1.y * v2.z - v1.z * v2.y;
    const double cy = -(v1.x * v2.z - v1.z * v2.x);
    const double cz = v1.x * v2.y - v1.y * v2.x;
    double mag = cx * cx;
    mag += cy * cy;
    mag += cz * cz;
    mag = std::sqrt(mag);
    const double dot = v1.x * v2.x + v1.y * v2.y + v1.z * v2.z;
    return std::atan2(mag, dot);
}

// Direct (cos lat cos lon, cos lat sin lon, sin lat) unit vector.  mathRoutines.h:172-178,318-331
inline Vec3 unit_from_lat(LatLon p) {
    Vec3 v;
    v.x = std::cos(p.lat) * std::cos(p.lon);
    v.y = std::cos(p.lat) * std::sin(p.lon);
    v.z = std::sin(p.lat);
    return v;
}

// Normalised chord midpoint; lon in (-pi, pi].  mathRoutines.h:167-186
inline LatLon chord_midpoint(LatLon a, LatLon b) {
    const Vec3 p = unit_from_lat(a), q = unit_from_lat(b);
    const double cx = 0.5 * (p.x + q.x), cy = 0.5 * (p.y + q.y), cz = 0.5 * (p.z + q.z);
    LatLon c;
    c.lat = std::atan2(cz, std::sqrt(cx * cx + cy * cy));
    c.lon = std::atan2(cy, cx);
    return c;
}

// Unit normal (in the map centred on the chord midpoint, unit sphere) of the segment a->b,
// rotated +90deg from the tangent.  mathRoutines.h:208-242
inline Vec2 edge_normal_in_map(LatLon a, LatLon b) {
    const LatLon c = chord_midpoint(a, b);
    const double r = 1.0;
    double m = map_factor(c, a);
    const double c1x = m * r * (std::cos(a.lat) * std::sin(a.lon - c.lon));
    const double c1y = m * r * (std::sin(a.lat) * std::cos(c.lat) - std::cos(a.lat) * std::sin(c.lat) * std::cos(a.lon - c.lon));
    m = map_factor(c, b);
    const double c2x = m * r * (std::cos(b.lat) * std::sin(b.lon - c.lon));
    const double c2y = m * r * (std::sin(b.lat) * std::cos(c.lat) - std::cos(b.lat) * std::sin(c.lat) * std::cos(b.lon - c.lon));
    double xx = c2x - c1x, yy = c2y - c1y;
    const double mag = std::sqrt(xx * xx + yy * yy);
    xx /= mag;
    yy /= mag;
    return Vec2{-yy, xx};
}

// Unit normal of the great-circle plane through a and b (a x b normalised).  mathRoutines.h:247-266
inline Vec3 great_circle_normal(LatLon a, LatLon b) {
    const Vec3 c1 = unit_from_colat(a), c2 = unit_from_colat(b);
    Vec3 n;
    n.x = c1.y * c2.z - c1.z * c2.y;
    n.y = -(c1.x * c2.z - c1.z * c2.x);
    n.z = c1.x * c2.y - c1.y * c2.x;
    const double mag = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    n.x /= mag;
    n.y /= mag;
    n.z /= mag;
    return n;
}

// Intersection of great circles (p1,p2) and (p3,p4); lon wrapped to [0, 2pi).  mathRoutines.h:303-369
inline LatLon great_circle_intersection(LatLon p1, LatLon p2, LatLon p3, LatLon p4) {
    const Vec3 c1 = unit_from_lat(p1), c2 = unit_from_lat(p2), c3 = unit_from_lat(p3), c4 = unit_from_lat(p4);
    double n1x = c1.y * c2.z - c1.z * c2.y;
    double n1y = -(c1.x * c2.z - c1.z * c2.x);
    double n1z = c1.x * c2.y - c1.y * c2.x;
    double n2x = c3.y * c4.z - c3.z * c4.y;
    double n2y = -(c3.x * c4.z - c3.z * c4.x);
    double n2z = c3.x * c4.y - c3.y * c4.x;
    double mag1 = std::sqrt(n1x * n1x + n1y * n1y + n1z * n1z);
    const double mag2 = std::sqrt(n2x * n2x + n2y * n2y + n2z * n2z);
    n1x /= mag1; n1y /= mag1; n1z /= mag1;
    n2x /= mag2; n2y /= mag2; n2z /= mag2;
    double mx = n1y * n2z - n2y * n1z;
    double my = n2x * n1z - n1x * n2z;
    double mz = n1x * n2y - n2x * n1y;
    mag1 = std::sqrt(mx * mx + my * my + mz * mz);
    mx /= mag1; my /= mag1; mz /= mag1;
    LatLon s;
    s.lat = std::asin(mz / 1.0);
    s.lon = std::atan2(my, mx);
    if (s.lon < 0.0) s.lon += 2 * kPi;
    return s;
}

// Planar triangle area with one corner (xc,yc).  mathRoutines.h:403-407
inline double planar_triangle_area(double xc, double yc, double x1, double x2, double y1, double y2) {
    return 0.5 * std::fabs(xc * (y1 - y2) + x1 * (y2 - yc) + x2 * (yc - y1));
}

// Spherical triangle area by spherical excess from three acos sides.  mathRoutines.h:410-429
inline double spherical_triangle_area(LatLon p1, LatLon p2, LatLon p3, double r) {
    const double c = std::fabs(std::acos(std::sin(p1.lat) * std::sin(p2.lat) + std::cos(p1.lat) * std::cos(p2.lat) * std::cos(std::fabs(p2.lon - p1.lon))));
    const double a = std::fabs(std::acos(std::sin(p2.lat) * std::sin(p3.lat) + std::cos(p2.lat) * std::cos(p3.lat) * std::cos(std::fabs(p3.lon - p2.lon))));
    const double b = std::fabs(std::acos(std::sin(p3.lat) * std::sin(p1.lat) + std::cos(p3.lat) * std::cos(p1.lat) * std::cos(std::fabs(p1.lon - p3.lon))));
    const double A = std::fabs(std::acos((std::cos(a) - std::cos(b) * std::cos(c)) / (std::sin(b) * std::sin(c))));
    const double B = std::fabs(std::acos((std::cos(b) - std::cos(a) * std::cos(c)) / (std::sin(a) * std::sin(c))));
    const double C = std::fabs(std::acos((std::cos(c) - std::cos(b) * std::cos(a)) / (std::sin(b) * std::sin(a))));
    return (r * r) * std::fabs((A + B + C) - kPi);
}

// Signed angle, in the unit-sphere map centred on c, from p2 to p1, folded into [-pi, pi].
// mathRoutines.h:372-397 (the atan2(det,dot) value computed there is overwritten; only the
// difference of the two polar angles survives).
inline double map_angle_between(LatLon p1, LatLon p2, LatLon c) {
    const Vec2 a = map_project(c, p1, 1.0), b = map_project(c, p2, 1.0);
    const double ang1 = std::atan2(a.y - 0.0, a.x - 0.0);
    const double ang2 = std::atan2(b.y - 0.0, b.x - 0.0);
    double angle = ang1 - ang2;
    if (angle > kPi) angle -= 2 * kPi;
    else if (angle < -kPi) angle += 2 * kPi;
    return angle;
}

}  // namespace odis
