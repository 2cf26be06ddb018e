/* Build shim for oracle/_ref (see block.h): a logger that swallows everything. */
#ifndef ORACLE_SHIM_GR_LOGGER_H
#define ORACLE_SHIM_GR_LOGGER_H
namespace gr {
class logger
{
public:
    template <class... A> void debug(A&&...) {}
    template <class... A> void info(A&&...) {}
    template <class... A> void warn(A&&...) {}
    template <class... A> void error(A&&...) {}
};
} // namespace gr
#define GR_LOG_DEBUG(...) ((void)0)
#endif
