// Per-thread error text shared by the C ABI translation units (odis_last_error()).
#pragma once
#include <string>
namespace odis {
extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);   // stores msg, returns code
}  // namespace odis
