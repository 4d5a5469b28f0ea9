/*
 * ref_stubs.cpp -- TEST INFRASTRUCTURE (oracle/_ref).
 * Link-time stand-ins for the four cage symbols whose own translation units need absent
 * third-party code (plf_list, zlib, avir).  None is reachable from the hot path; tasksRunBlocking
 * runs the invocations serially so that meshChunking(parallelize) would still be correct.
 */
#include <cage-core/files.h>
#include <cage-core/tasks.h>

namespace cage
{
	namespace privat
	{
		void tasksRunBlocking(TaskCreateConfig &&task)
		{
			for (uint32 i = 0; i < task.invocations; i++)
				task.runner(task, i);
		}
	}

	namespace detail
	{
		void imageResize(const float *, uint32, uint32, uint32, float *, uint32, uint32, uint32, uint32)
		{
			CAGE_THROW_CRITICAL(Exception, "imageResize is not part of the oracle build");
		}
	}

	Holder<File> archiveOpenZipTry(Holder<File> &&)
	{
		return {};
	}

	void archiveCreateZip(const String &, const String &)
	{
		CAGE_THROW_CRITICAL(Exception, "zip archives are not part of the oracle build");
	}
}
