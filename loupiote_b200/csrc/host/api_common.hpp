// Shared helpers for the C ABI translation units.
#pragma once
#include <exception>
#include <new>
#include <string>

#include "loupiote.h"

namespace lp {
struct Scene;
extern thread_local std::string g_last_error;
lp_status fail(lp_status st, const std::string &msg);
Scene &scene_of(lp_scene *s);
}  // namespace lp

// No exception may cross the ABI.
#define LP_TRY(...)                                          \
  try {                                                      \
    __VA_ARGS__                                              \
  } catch (const std::bad_alloc &) {                         \
    return lp::fail(LP_ERR_OOM, "out of host memory");       \
  } catch (const std::exception &e) {                        \
    return lp::fail(LP_ERR_INVALID_ARG, e.what());           \
  }

// Function-try-block tail for the C ABI entry points of the CUDA translation units:
//   LP_API lp_status f(...) try { ... } LP_ABI_CATCH
// std::bad_alloc and friends from std::vector / std::string inside an entry point become status
// codes like everywhere else.
#define LP_ABI_CATCH                                         \
  catch (const std::bad_alloc &) {                           \
    return lp::fail(LP_ERR_OOM, "out of host memory");       \
  }                                                          \
  catch (const std::exception &e) {                          \
    return lp::fail(LP_ERR_INVALID_ARG, e.what());           \
  }
