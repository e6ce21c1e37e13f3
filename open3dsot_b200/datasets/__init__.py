"""Data side of the path.  `get_dataset` mirrors datasets/__init__.py:8-46 for what is implemented here: the KITTI reader
and, on top of it, the device-side batch construction (`train_siamese`) or the raw tracklets (`test`)."""


def get_dataset(config, type='train', **kwargs):
    if config.dataset != 'kitti':
        raise NotImplementedError(f"dataset '{config.dataset}': only the KITTI reader is implemented "
                                  "(nuScenes / Waymo need nuscenes-devkit / the Waymo converter; DESIGN.md section 9)")
    from .kitti import kittiDataset
    data = kittiDataset(path=config.path, split=kwargs.get('split', 'train'), category_name=config.category_name,
                        coordinate_mode=config.coordinate_mode, preloading=config.preloading,
                        preload_offset=config.preload_offset if type != 'test' else -1)
    if type == 'train_siamese':
        from .device_sampler import DeviceSiameseSampler
        return DeviceSiameseSampler(data.tracklets(), config, kwargs.get('device', 'cuda'))
    if type.lower() == 'train_motion':
        from .device_sampler import DeviceMotionSampler
        return DeviceMotionSampler(data.tracklets(), config, kwargs.get('device', 'cuda'))
    return data.tracklets()
