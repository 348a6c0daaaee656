// empty stand-in: the reference includes this Boost header but uses nothing from it on the file-sink path
#pragma once
