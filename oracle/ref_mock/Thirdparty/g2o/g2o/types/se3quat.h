// ORACLE — TEST INFRASTRUCTURE ONLY.  Empty stand-in: the hot path never touches g2o (Common.h includes it for the back end).
#pragma once
