// Empty stand-in: the reference includes <boost/container/set.hpp> (include/index_bipartite.h:1) but never uses it.
#pragma once
