// pointCloudMeasure.hpp -- empty stand-in: included by Main_KSS_ICP.cpp:29, never called.
#pragma once
