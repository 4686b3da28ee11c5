// Same include path as the reference; everything lives in regressors.hpp.
#include <sequential-line-search/regressors.hpp>
