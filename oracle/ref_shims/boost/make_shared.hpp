#include <boost/shared_ptr.hpp>
