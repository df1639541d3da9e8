// TEST INFRASTRUCTURE ONLY: numpy::initialize() stand-in lives in ../python.hpp
#pragma once
#include "../python.hpp"
