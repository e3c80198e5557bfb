// TEST INFRASTRUCTURE -- forwards to the stand-ins (tests/cxx/aligator_stub/aligator_standins.hpp).
#pragma once
#include "aligator_standins.hpp"
