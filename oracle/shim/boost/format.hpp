/* Build shim for oracle/_ref: the reference includes boost/format.hpp without using it on this path. */
