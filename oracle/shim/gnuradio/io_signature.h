/* Build shim for oracle/_ref (see block.h). */
#ifndef ORACLE_SHIM_GR_IO_SIGNATURE_H
#define ORACLE_SHIM_GR_IO_SIGNATURE_H
#include <memory>
namespace gr {
class io_signature
{
public:
    typedef std::shared_ptr<io_signature> sptr;
    static sptr make(int, int, int) { return std::make_shared<io_signature>(); }
};
} // namespace gr
#endif
