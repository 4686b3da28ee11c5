// Same include path as the reference; everything lives in acquisition.hpp.
#include <sequential-line-search/acquisition.hpp>
