// ORACLE SCAFFOLDING: read-only mmap wrapper standing in for boost::iostreams::mapped_file_source.
#pragma once
#include <fcntl.h>
#include <stdexcept>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
namespace boost { namespace iostreams {
class mapped_file_source {
    const char* d_ = nullptr; size_t n_ = 0; bool open_ = false;
   public:
    mapped_file_source() = default;
    ~mapped_file_source() { close(); }
    void open(const std::string& path) {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("mapped_file_source: cannot open " + path);
        struct stat st; if (fstat(fd, &st) != 0) { ::close(fd); throw std::runtime_error("fstat"); }
        n_ = static_cast<size_t>(st.st_size);
        if (n_ > 0) {
            void* p = mmap(nullptr, n_, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { ::close(fd); throw std::runtime_error("mmap"); }
            d_ = static_cast<const char*>(p);
        }
        ::close(fd); open_ = true;
    }
    bool is_open() const { return open_; }
    const char* data() const { return d_; }
    size_t size() const { return n_; }
    void close() { if (d_) munmap(const_cast<char*>(d_), n_); d_ = nullptr; n_ = 0; open_ = false; }
};
}}  // namespace boost::iostreams
