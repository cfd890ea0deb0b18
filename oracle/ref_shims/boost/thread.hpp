#ifndef REF_SHIM_BOOST_THREAD
#define REF_SHIM_BOOST_THREAD
#include <boost/thread/mutex.hpp>
#include <boost/thread/once.hpp>
#include <boost/bind.hpp>
#include <memory>
#include <thread>
#include <vector>
/* boost::thread_group / this_thread::disable_interruption for HomotopyClassPlanner::optimizeAllTEBs (TEST INFRASTRUCTURE) */
namespace boost {
typedef std::thread thread;
class thread_group {
 public:
  template <class F> thread* create_thread(F f) { ts_.emplace_back(new std::thread(f)); return ts_.back().get(); }
  void join_all() { for (auto& t : ts_) if (t->joinable()) t->join(); }
  ~thread_group() { join_all(); }
 private:
  std::vector<std::unique_ptr<std::thread> > ts_;
};
namespace this_thread { struct disable_interruption {}; }
}
#endif
