// TEST INFRASTRUCTURE — nothing of highgui is used by the filter sources.
#pragma once
#include "opencv2/core.hpp"
