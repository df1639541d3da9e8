// TEST INFRASTRUCTURE ONLY: empty stand-in (see python.hpp)
#pragma once
