// "<double><h|d|m|y>" -> earth days (reference src/utility/TimeConverter.cpp:4-48; tests/TimeConverterTest.cpp:4-47).
#pragma once
#include <string>

class TimeConverter {
public:
    static double convertToEarthDays(std::string &time);
};
