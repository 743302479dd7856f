// internal to the mini MEX runtime: what the harness needs beyond the public mex.h
#pragma once
#include <stddef.h>
#include <string>

namespace minimex {
struct MexError {
  std::string id, msg;
};
void run_atexit();        // what `clear mex` / MATLAB exit does with the registered handlers
size_t n_atexit();
}  // namespace minimex
