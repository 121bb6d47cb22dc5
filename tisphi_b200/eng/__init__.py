"""Host-side mirror of tiSPHi's ``eng`` package: same module, class and method names, bodies re-pointed to the
B200 engine (libtisphi_b200.so through tisphi_b200._lib).  No Taichi, no CPU fallback."""
