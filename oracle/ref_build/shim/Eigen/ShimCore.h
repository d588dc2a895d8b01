// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for the subset of the Eigen API that the GeodesicODIS
// reference sources touch, written from scratch so that the UNMODIFIED
// reference .cpp files under /root/reference/src can be compiled in an image
// that has no Eigen (Eigen is an un-vendored, un-pinned header dependency of the
// reference; the Makefile only adds -I/usr/local/, /root/reference/Makefile:42).
//
// Semantics restated from Eigen 3.3/3.4's published behaviour, only where the
// reference's arithmetic depends on it:
//   * SparseMatrix<double,RowMajor>::setFromTriplets: CSR, inner indices
//     ascending, duplicates summed in input order.
//   * sparse(row-major) * dense-vector: per row, tmp = 0; tmp += a_ij * x_j in
//     ascending j; res_i += 1.0 * tmp   (res zero-initialised by "dst = product").
//   * (scalar * sparse) * vector: the scalar multiplies each coefficient before
//     the coefficient multiplies x_j.
//   * dst = P1 + P2 (two products): dst = P1; dst += P2.
//   * sparse * sparse: row-wise accumulation, contributions to one output
//     coefficient added in ascending inner index k; exact zeros are kept until
//     prune().
//   * dynamic dense inverse(): partial-pivot LU (unblocked right-looking, what
//     Eigen's PartialPivLU does below its blocking threshold), then the solve
//     against the permuted identity. Only the reference's advection / RBF
//     operators (out of the linear hot path) depend on this.
// Everything else (FullPivLU, LLT, SimplicialLDLT) is constructed by the
// reference but its result is never read at HEAD, so compute() is a no-op.
#ifndef ODIS_ORACLE_EIGEN_SHIM_CORE_H
#define ODIS_ORACLE_EIGEN_SHIM_CORE_H

#include <vector>
#include <array>
#include <cmath>
#include <cstddef>
#include <algorithm>
#include <utility>

namespace Eigen {

const int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1 };

// ---------------------------------------------------------------- dense ----
template <typename S>
struct DenseData {
    int r = 0, c = 0;
    std::vector<S> v;                       // row-major, always
};

struct IdentityExpr;

template <typename S, int R, int C, int Opt = 0>
class Matrix {
public:
    DenseData<S> d;
    Matrix() { d.r = (R > 0 ? R : 0); d.c = (C > 0 ? C : 0); d.v.assign((size_t)d.r * d.c, S(0)); }
    explicit Matrix(int n) {
        // vector-like construction: VectorXd(n) or row-vector(n)
        if (C == 1) { d.r = n; d.c = 1; } else if (R == 1) { d.r = 1; d.c = n; } else { d.r = n; d.c = n; }
        d.v.assign((size_t)d.r * d.c, S(0));
    }
    Matrix(int r, int c) { d.r = r; d.c = c; d.v.assign((size_t)r * c, S(0)); }
    template <int R2, int C2, int O2>
    Matrix(const Matrix<S, R2, C2, O2>& o) { d = o.d; }
    template <int R2, int C2, int O2>
    Matrix& operator=(const Matrix<S, R2, C2, O2>& o) { d = o.d; return *this; }
    template <typename VR, typename = decltype(std::declval<VR>().v)>
    Matrix& operator=(const VR& r) { d.r = (int)r.v.size(); d.c = 1; d.v.assign(r.v.begin(), r.v.end()); return *this; }

    int rows() const { return d.r; }
    int cols() const { return d.c; }
    int size() const { return d.r * d.c; }
    S* data() { return d.v.data(); }
    const S* data() const { return d.v.data(); }

    S& operator()(int i, int j) { return d.v[(size_t)i * d.c + j]; }
    S operator()(int i, int j) const { return d.v[(size_t)i * d.c + j]; }
    S& operator()(int i) { return d.v[i]; }
    S operator()(int i) const { return d.v[i]; }
    S& operator[](int i) { return d.v[i]; }
    S operator[](int i) const { return d.v[i]; }
    S coeff(int i, int j) const { return d.v[(size_t)i * d.c + j]; }
    S coeff(int i) const { return d.v[i]; }
    S& coeffRef(int i, int j) { return d.v[(size_t)i * d.c + j]; }
    void setZero() { std::fill(d.v.begin(), d.v.end(), S(0)); }
    void resize(int r, int c) { d.r = r; d.c = c; d.v.assign((size_t)r * c, S(0)); }

    static IdentityExpr Identity(int r, int c);

    // Partial-pivot LU inverse (see header comment).
    Matrix<S, Dynamic, Dynamic, 0> inverse() const {
        const int n = d.r;
        std::vector<S> lu(d.v);                 // row-major n x n
        std::vector<int> piv(n);
        for (int i = 0; i < n; i++) piv[i] = i;
        auto at = [&](int i, int j) -> S& { return lu[(size_t)i * n + j]; };
        for (int k = 0; k < n; k++) {
            int p = k; S best = std::fabs(at(k, k));
            for (int i = k + 1; i < n; i++) {
                S a = std::fabs(at(i, k));
                if (a > best) { best = a; p = i; }
            }
            if (p != k) {
                for (int j = 0; j < n; j++) std::swap(at(k, j), at(p, j));
                std::swap(piv[k], piv[p]);
            }
            if (at(k, k) != S(0)) {
                for (int i = k + 1; i < n; i++) at(i, k) /= at(k, k);
            }
            for (int j = k + 1; j < n; j++)          // column-major sweep of the rank-1 update
                for (int i = k + 1; i < n; i++) at(i, j) -= at(i, k) * at(k, j);
        }
        Matrix<S, Dynamic, Dynamic, 0> inv(n, n);
        for (int col = 0; col < n; col++) {
            std::vector<S> x(n);
            for (int i = 0; i < n; i++) x[i] = (piv[i] == col) ? S(1) : S(0);   // P * e_col
            for (int j = 0; j < n; j++)                                          // L y = Pb (unit lower)
                for (int i = j + 1; i < n; i++) x[i] -= at(i, j) * x[j];
            for (int j = n - 1; j >= 0; j--) {                                   // U x = y
                x[j] /= at(j, j);
                for (int i = 0; i < j; i++) x[i] -= at(i, j) * x[j];
            }
            for (int i = 0; i < n; i++) inv(i, col) = x[i];
        }
        return inv;
    }
};

typedef Matrix<double, Dynamic, Dynamic, 0> MatrixXd;
typedef Matrix<double, Dynamic, 1, 0> VectorXd;

// ------------------------------------------------- vector expressions ------
// A materialised vector result (what a product or a sum evaluates to).
struct VecResult {
    std::vector<double> v;
    int size() const { return (int)v.size(); }
};
struct ScaledVecRef {            // scalar * Map  (lazy)
    double s; const double* p; int n;
};

template <typename T> class Map;

template <int R, int C, int O>
class Map<Matrix<double, R, C, O>> {
public:
    double* p; int n;
    Map(double* ptr, int len) : p(ptr), n(len) {}
    int size() const { return n; }
    double& operator()(int i) { return p[i]; }
    double operator()(int i) const { return p[i]; }
    double coeff(int i) const { return p[i]; }
    Map& operator=(const VecResult& r) { for (int i = 0; i < n; i++) p[i] = r.v[i]; return *this; }
    Map& operator=(const Map& o) { for (int i = 0; i < n; i++) p[i] = o.p[i]; return *this; }
    Map& operator+=(const VecResult& r) { for (int i = 0; i < n; i++) p[i] += r.v[i]; return *this; }
    Map& operator+=(const ScaledVecRef& r) { for (int i = 0; i < n; i++) p[i] += r.s * r.p[i]; return *this; }
    Map& operator*=(double s) { for (int i = 0; i < n; i++) p[i] *= s; return *this; }
};

template <int R, int C, int O>
inline ScaledVecRef operator*(double s, const Map<Matrix<double, R, C, O>>& m) { return ScaledVecRef{s, m.p, m.n}; }

// dst = P1 + P2  ==  dst = P1; dst += P2   (element-wise t1 + t2)
inline VecResult operator+(const VecResult& a, const VecResult& b) {
    VecResult r; r.v.resize(a.v.size());
    for (size_t i = 0; i < a.v.size(); i++) r.v[i] = a.v[i] + b.v[i];
    return r;
}

// ---------------------------------------------------------------- sparse ---
template <typename S>
class Triplet {
public:
    Triplet() : r_(0), c_(0), v_(S(0)) {}
    Triplet(int r, int c, S v) : r_(r), c_(c), v_(v) {}
    int row() const { return r_; }
    int col() const { return c_; }
    S value() const { return v_; }
    S& valueRef() { return v_; }
private:
    int r_, c_; S v_;
};

template <typename S, int Opt> class SparseMatrix;

template <typename S, int Opt>
struct ScaledSparse {                     // scalar * sparse (lazy)
    S s; const SparseMatrix<S, Opt>* m;
};

template <typename S, int Opt = 0>
class SparseMatrix {
public:
    int nr = 0, nc = 0;
    mutable std::vector<int> ptr;         // CSR row pointers (nr+1)
    mutable std::vector<int> idx;
    mutable std::vector<S> val;
    mutable std::vector<Triplet<S>> pending;   // insert()ed, not yet merged

    SparseMatrix() { ptr.assign(1, 0); }
    SparseMatrix(int r, int c) : nr(r), nc(c) { ptr.assign((size_t)r + 1, 0); }
    SparseMatrix(const ScaledSparse<S, Opt>& e) { *this = e; }

    int rows() const { return nr; }
    int cols() const { return nc; }
    int nonZeros() const { flush(); return (int)idx.size(); }
    template <typename T> void reserve(const T&) {}
    void makeCompressed() { flush(); }
    // compressed-storage accessors of Eigen::SparseMatrix (what integration/*.cpp hands to the C ABI)
    const int* outerIndexPtr() const { flush(); return ptr.data(); }
    const int* innerIndexPtr() const { flush(); return idx.data(); }
    const S* valuePtr() const { flush(); return val.data(); }

    void buildFrom(std::vector<Triplet<S>>& t) {
        // stable by (row, col); duplicates summed in input order
        std::stable_sort(t.begin(), t.end(), [](const Triplet<S>& a, const Triplet<S>& b) {
            if (a.row() != b.row()) return a.row() < b.row();
            return a.col() < b.col();
        });
        ptr.assign((size_t)nr + 1, 0); idx.clear(); val.clear();
        idx.reserve(t.size()); val.reserve(t.size());
        int lastr = -1, lastc = -1;
        for (const auto& e : t) {
            if (e.row() == lastr && e.col() == lastc) { val.back() += e.value(); continue; }
            idx.push_back(e.col()); val.push_back(e.value());
            ptr[(size_t)e.row() + 1]++;
            lastr = e.row(); lastc = e.col();
        }
        for (int i = 0; i < nr; i++) ptr[(size_t)i + 1] += ptr[i];
    }
    template <typename It>
    void setFromTriplets(It b, It e) {
        std::vector<Triplet<S>> t(b, e);
        pending.clear();
        buildFrom(t);
    }
    S& insert(int i, int j) {
        pending.push_back(Triplet<S>(i, j, S(0)));
        return pending.back().valueRef();
    }
    void flush() const {
        if (pending.empty()) return;
        std::vector<Triplet<S>> t;
        t.reserve(idx.size() + pending.size());
        for (int i = 0; i < nr; i++)
            for (int k = ptr[i]; k < ptr[(size_t)i + 1]; k++) t.push_back(Triplet<S>(i, idx[k], val[k]));
        for (auto& e : pending) t.push_back(e);
        pending.clear();
        const_cast<SparseMatrix*>(this)->buildFrom(t);
    }
    void prune(S ref) {
        flush();
        std::vector<int> np((size_t)nr + 1, 0), ni; std::vector<S> nv;
        for (int i = 0; i < nr; i++) {
            for (int k = ptr[i]; k < ptr[(size_t)i + 1]; k++)
                if (std::fabs(val[k]) > ref) { ni.push_back(idx[k]); nv.push_back(val[k]); }
            np[(size_t)i + 1] = (int)ni.size();
        }
        ptr.swap(np); idx.swap(ni); val.swap(nv);
    }
    SparseMatrix& operator=(const ScaledSparse<S, Opt>& e) {
        e.m->flush();
        nr = e.m->nr; nc = e.m->nc; ptr = e.m->ptr; idx = e.m->idx; val = e.m->val; pending.clear();
        for (auto& x : val) x = e.s * x;
        return *this;
    }
    SparseMatrix operator-() const {
        flush();
        SparseMatrix r(*this);
        for (auto& x : r.val) x = -x;
        return r;
    }
    SparseMatrix& operator+=(const SparseMatrix& o) {
        flush(); o.flush();
        std::vector<int> np((size_t)nr + 1, 0), ni; std::vector<S> nv;
        for (int i = 0; i < nr; i++) {
            int a = ptr[i], ae = ptr[(size_t)i + 1], b = o.ptr[i], be = o.ptr[(size_t)i + 1];
            while (a < ae || b < be) {
                if (b >= be || (a < ae && idx[a] < o.idx[b])) { ni.push_back(idx[a]); nv.push_back(val[a]); a++; }
                else if (a >= ae || o.idx[b] < idx[a]) { ni.push_back(o.idx[b]); nv.push_back(o.val[b]); b++; }
                else { ni.push_back(idx[a]); nv.push_back(val[a] + o.val[b]); a++; b++; }
            }
            np[(size_t)i + 1] = (int)ni.size();
        }
        ptr.swap(np); idx.swap(ni); val.swap(nv);
        return *this;
    }
    S coeff(int i, int j) const {
        flush();
        for (int k = ptr[i]; k < ptr[(size_t)i + 1]; k++) if (idx[k] == j) return val[k];
        return S(0);
    }
};

template <typename S, int Opt>
inline ScaledSparse<S, Opt> operator*(double s, const SparseMatrix<S, Opt>& m) { return ScaledSparse<S, Opt>{(S)s, &m}; }
template <typename S, int Opt>
inline ScaledSparse<S, Opt> operator*(double s, const ScaledSparse<S, Opt>& m) { return ScaledSparse<S, Opt>{(S)(s * m.s), m.m}; }

// sparse * vector kernels ----------------------------------------------------
template <typename S, int Opt>
inline VecResult spmv(const SparseMatrix<S, Opt>& A, const double* x) {
    A.flush();
    VecResult r; r.v.assign((size_t)A.nr, 0.0);
    // Eigen runs row-major sparse * dense products row-parallel when built with OpenMP (rows are independent,
    // so results do not depend on the thread count); only the OPENMP=1 variant of the Makefile enables it
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < A.nr; i++) {
        double tmp = 0;
        for (int k = A.ptr[i]; k < A.ptr[(size_t)i + 1]; k++) tmp += A.val[k] * x[A.idx[k]];
        r.v[i] += 1.0 * tmp;
    }
    return r;
}
template <typename S, int Opt>
inline VecResult spmv_scaled(S s, const SparseMatrix<S, Opt>& A, const double* x) {
    A.flush();
    VecResult r; r.v.assign((size_t)A.nr, 0.0);
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < A.nr; i++) {
        double tmp = 0;
        for (int k = A.ptr[i]; k < A.ptr[(size_t)i + 1]; k++) tmp += (s * A.val[k]) * x[A.idx[k]];
        r.v[i] += 1.0 * tmp;
    }
    return r;
}
template <typename S, int Opt, int R, int C, int O>
inline VecResult operator*(const SparseMatrix<S, Opt>& A, const Map<Matrix<double, R, C, O>>& x) { return spmv(A, x.p); }
template <typename S, int Opt, int R, int C, int O>
inline VecResult operator*(const ScaledSparse<S, Opt>& A, const Map<Matrix<double, R, C, O>>& x) { return spmv_scaled(A.s, *A.m, x.p); }
template <typename S, int Opt, int R, int C, int O>
inline VecResult operator*(const SparseMatrix<S, Opt>& A, const Matrix<double, R, C, O>& x) { return spmv(A, x.data()); }

// sparse * sparse -------------------------------------------------------------
template <typename S, int Opt>
inline SparseMatrix<S, Opt> operator*(const SparseMatrix<S, Opt>& A, const SparseMatrix<S, Opt>& B) {
    A.flush(); B.flush();
    SparseMatrix<S, Opt> Cm(A.nr, B.nc);
    std::vector<S> acc((size_t)B.nc, S(0));
    std::vector<char> mark((size_t)B.nc, 0);
    std::vector<int> touched;
    for (int i = 0; i < A.nr; i++) {
        touched.clear();
        for (int ka = A.ptr[i]; ka < A.ptr[(size_t)i + 1]; ka++) {
            const int k = A.idx[ka]; const S a = A.val[ka];
            for (int kb = B.ptr[k]; kb < B.ptr[(size_t)k + 1]; kb++) {
                const int j = B.idx[kb];
                if (!mark[j]) { mark[j] = 1; acc[j] = a * B.val[kb]; touched.push_back(j); }
                else acc[j] += a * B.val[kb];
            }
        }
        std::sort(touched.begin(), touched.end());
        for (int j : touched) { Cm.idx.push_back(j); Cm.val.push_back(acc[j]); mark[j] = 0; }
        Cm.ptr[(size_t)i + 1] = (int)Cm.idx.size();
    }
    return Cm;
}
template <typename S, int Opt>
inline SparseMatrix<S, Opt> operator*(const ScaledSparse<S, Opt>& A, const SparseMatrix<S, Opt>& B) {
    SparseMatrix<S, Opt> As(A);
    return As * B;
}

// Identity(n,n).sparseView()
struct IdentityExpr {
    int r, c;
    SparseMatrix<double, RowMajor> sparseView() const {
        SparseMatrix<double, RowMajor> I(r, c);
        const int n = std::min(r, c);
        I.idx.resize(n); I.val.assign(n, 1.0);
        for (int i = 0; i < n; i++) I.idx[i] = i;
        for (int i = 0; i < r; i++) I.ptr[(size_t)i + 1] = std::min(i + 1, n);
        return I;
    }
};
template <typename S, int R, int C, int Opt>
inline IdentityExpr Matrix<S, R, C, Opt>::Identity(int r, int c) { return IdentityExpr{r, c}; }

// ------------------------------------------------- inert factorisations ----
template <typename M> class FullPivLU { public: template <typename T> FullPivLU& compute(const T&) { return *this; } };
template <typename M> class LLT { public: template <typename T> LLT& compute(const T&) { return *this; } };
template <typename M> class SimplicialLDLT { public: template <typename T> SimplicialLDLT& compute(const T&) { return *this; } };
template <typename M> class ColPivHouseholderQR { public: template <typename T> ColPivHouseholderQR& compute(const T&) { return *this; } };

}  // namespace Eigen

#endif
