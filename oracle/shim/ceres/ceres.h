// ceres/ceres.h -- TEST INFRASTRUCTURE ONLY (oracle/).  3d/carto_math.h names ceres::atan2 in a
// template the scan-match path never instantiates.
#ifndef GLOC_ORACLE_CERES_SHIM_H_
#define GLOC_ORACLE_CERES_SHIM_H_
#include <cmath>
namespace ceres { using std::atan2; }
#endif
