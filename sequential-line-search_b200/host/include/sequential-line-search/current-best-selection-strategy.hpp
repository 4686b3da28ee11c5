// Same include path as the reference; everything lives in optimizers.hpp.
#include <sequential-line-search/optimizers.hpp>
