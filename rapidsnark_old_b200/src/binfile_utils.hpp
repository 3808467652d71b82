// iden3 "binfile" container reader - same interface as the reference's BinFileUtils
// (src/binfile_utils.hpp:9-46, src/binfile_utils.cpp:14-147): magic, version, nSections, then
// {u32 type, u64 size, payload} sections.  Differences by design: the file is mmap'ed read-only and used
// in place (the reference mmaps, copies the whole file to malloc memory and unmaps - 2x the zkey in RAM,
// binfile_utils.cpp:28-32), and exceptions are thrown by value, with the reference's messages.
#ifndef B200_BINFILE_UTILS_HPP
#define B200_BINFILE_UTILS_HPP
#include <stdint.h>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace BinFileUtils {

class BinFile {
    void *addr;
    uint64_t size;
    uint64_t pos;
    struct Section { void *start; uint64_t size; };
    std::map<int, std::vector<Section>> sections;
    std::string type;
    uint32_t version;
    Section *readingSection;

public:
    BinFile(std::string fileName, std::string type, uint32_t maxVersion);
    ~BinFile();
    BinFile(const BinFile &) = delete;
    BinFile &operator=(const BinFile &) = delete;

    void startReadSection(uint32_t sectionId, uint32_t sectionPos = 0);
    void endReadSection(bool check = true);
    void *getSectionData(uint32_t sectionId, uint32_t sectionPos = 0);
    uint64_t getSectionSize(uint32_t sectionId, uint32_t sectionPos = 0);
    uint32_t readU32LE();
    uint64_t readU64LE();
    void *read(uint64_t l);
};

std::unique_ptr<BinFile> openExisting(std::string filename, std::string type, uint32_t maxVersion);

// Scoped sequential reader of one section (startReadSection ... endReadSection with the size check).
class SectionReader {
    BinFile *f;
    bool open;

public:
    SectionReader(BinFile *file, uint32_t sectionId, uint32_t sectionPos = 0) : f(file), open(true) { f->startReadSection(sectionId, sectionPos); }
    ~SectionReader() { if (open) f->endReadSection(false); }
    uint32_t u32() { return f->readU32LE(); }
    uint64_t u64() { return f->readU64LE(); }
    void *raw(uint64_t len) { return f->read(len); }
    std::vector<uint8_t> bytes(uint64_t len) { const uint8_t *p = (const uint8_t *)f->read(len); return std::vector<uint8_t>(p, p + len); }
    void finish() { open = false; f->endReadSection(true); }
};

}  // namespace BinFileUtils
#endif
