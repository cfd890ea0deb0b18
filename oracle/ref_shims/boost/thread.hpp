#include <boost/thread/mutex.hpp>
#include <boost/thread/once.hpp>
