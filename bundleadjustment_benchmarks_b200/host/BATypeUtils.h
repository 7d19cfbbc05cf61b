// Scalar switch kept from the reference (src/BATypeUtils.h:6-7): edit the typedef (or build with
// -DBA_SCALAR_FLOAT) to switch the whole pipeline, GPU kernels included, between float and double.
// Eigen is not available in this environment, so the fixed-size Eigen typedefs of the reference are
// replaced by plain std::array/std::vector storage.
#ifndef BA_TYPE_UTILS_H
#define BA_TYPE_UTILS_H
#include <array>
#include <vector>

#ifdef BA_SCALAR_FLOAT
typedef float Scalar;
#else
//typedef float Scalar;
typedef double Scalar;
#endif

typedef std::array<Scalar, 2> Vector2X;
typedef std::array<Scalar, 3> Vector3X;
typedef std::array<Scalar, 9> Matrix3X;  // row-major 3x3
typedef std::vector<Scalar> VectorXX;

#endif
