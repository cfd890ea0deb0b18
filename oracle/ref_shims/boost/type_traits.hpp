#include <boost/utility.hpp>
