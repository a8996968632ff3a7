// qpOASES/Types.hpp stand-in (see ../qpOASES.hpp).  TEST INFRASTRUCTURE
#pragma once
namespace qpOASES {
typedef double real_t;
typedef int int_t;
enum PrintLevel { PL_NONE = 0 };
}  // namespace qpOASES
