// TEST INFRASTRUCTURE — boost::shared_ptr / make_shared as aliases of the std ones (see README.md).
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
}  // namespace boost
