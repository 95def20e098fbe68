// ORACLE — TEST INFRASTRUCTURE ONLY.  Empty stand-in: the hot path never touches Pangolin (viewer, out of scope).
#pragma once
